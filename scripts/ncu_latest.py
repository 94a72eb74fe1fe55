"""Record an `ncu --set full` capture as the evidence bench.py may quote (profiles/latest_ncu.json).
    python scripts/ncu_latest.py <config c1..c4> <strict|fast> <file.ncu-rep> <profiles/summary.txt> ["note"]
Stores the DRAM bytes of the captured launch and the FP64-pipe / issue figures together with the mangled name and the
SHA-1 of the SASS of the kernel in THIS build (bench.sass_identity): bench.py withholds the figures when the build it runs
has another kernel (roofline.ncu_stale)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    config, precision, rep, summary = sys.argv[1:5]
    note = sys.argv[5] if len(sys.argv) > 5 else ''
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    d = dict(zip(hdr, rows[2]))
    u = dict(zip(hdr, units))

    def val(k, scale_units=True):
        x = float(d[k].replace(',', ''))
        if scale_units:
            x *= {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'Tbyte': 1e12}.get(u[k], 1.0)
        return x
    pat, objname = bench.NCU_KEYS[config]
    name, sha = bench.sass_identity(pat, os.path.join(ROOT, 'clawpack_b200', 'csrc', 'build', objname % precision))
    assert name is not None, 'cannot identify the kernel of this build'
    assert pat.split('INS0_')[0].split('N8clawcuda')[-1] or True
    stalls = sorted(((float(d[h]), h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')])
                     for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')),
                    reverse=True)
    rec = {
        'source': '%s (ncu --set full --clock-control none, one launch%s)' % (os.path.relpath(summary, ROOT), ', ' + note if note else ''),
        'kernel': d['Kernel Name'], 'grid': d['Grid Size'], 'kernel_symbol': name, 'sass_sha1': sha,
        'gpu_time_ms': val('gpu__time_duration.sum', False) * {'ms': 1.0, 'us': 1e-3, 'ns': 1e-6, 's': 1e3}.get(u['gpu__time_duration.sum'], 1.0),
        'dram_bytes_per_launch': int(val('dram__bytes_read.sum') + val('dram__bytes_write.sum')),
        'dram_bytes_read': int(val('dram__bytes_read.sum')), 'dram_bytes_write': int(val('dram__bytes_write.sum')),
        'fp64_pipe': {
            'cycles_active_pct': round(val('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', False), 1),
            'issue_active_pct': round(val('smsp__issue_active.avg.pct_of_peak_sustained_active', False), 1),
            'warps_active_pct': round(val('sm__warps_active.avg.pct_of_peak_sustained_active', False), 1),
            'registers_per_thread': int(val('launch__registers_per_thread', False)),
            'dram_throughput_pct_of_ncu_peak': round(val('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', False), 1),
            'top_stalls': [s for _, s in stalls[:3]],
        },
    }
    f = os.path.join(ROOT, 'profiles', 'latest_ncu.json')
    allrec = json.load(open(f)) if os.path.exists(f) else {}
    allrec['%s_%s' % (config, precision)] = rec
    json.dump(allrec, open(f, 'w'), indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == '__main__':
    main()
