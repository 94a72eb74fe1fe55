"""Controller: the output-time loop around ``solver.evolve_to_time`` (pyclaw/src/pyclaw/controller.py:275-397).

Output styles 1-3, ``keep_copy`` frames, ``tfinal``, ``num_output_times``, frame files through ``Solution.write``
(``output_format='ascii'`` as in the reference, ``'binary'`` / ``'binary32'`` for raw dumps of large grids, ``None``
for no files), derived quantities (``compute_p`` -> ``<outdir>/_p/claw_p.*``), output functionals (``compute_F`` ->
``<outdir>/F.txt``), the ``overwrite`` guard, gauge files.  The plotting / xclaw-runner attributes of the reference
Controller are not part of this build."""
import copy
import logging
import os

import numpy as np


class Controller(object):
    def __init__(self):
        self.solution = None
        self.solver = None
        self.keep_copy = False
        self.frames = []
        self.write_aux_init = False
        self.write_aux_always = False
        self.output_format = 'ascii'                     # controller.py:141
        self.output_file_prefix = None
        self.output_options = {}
        self.outdir = './_output'
        self.overwrite = True
        self.logger = logging.getLogger('pyclaw.controller')
        self.tfinal = 1.0
        self.output_style = 1
        self.verbosity = 3
        self.num_output_times = 10
        self.out_times = np.linspace(0.0, self.tfinal, self.num_output_times)
        self.nstepout = 1
        self.start_frame = 0
        self.setplot = None
        self.file_prefix_p = 'claw_p'                    # controller.py:186
        self.compute_p = None                            # f(state): fills state.p (controller.py:188-190)
        self.compute_F = None                            # f(state): fills state.F (controller.py:192-195)
        self.F_file_name = 'F'

    @property
    def outdir_p(self):
        """Directory of the derived-quantity frames (controller.py:73-75)."""
        return os.path.join(self.outdir, '_p')

    @property
    def F_path(self):
        """File of the output functionals (controller.py:77-79)."""
        return os.path.join(self.outdir, self.F_file_name + '.txt')

    @property
    def verbosity(self):
        return self._verbosity

    @verbosity.setter
    def verbosity(self, value):
        self._verbosity = value

    def check_validity(self):
        if not self.solution.is_valid():
            raise Exception('The solution is not valid.')
        valid, reason = self.solver.is_valid()
        if not valid:
            raise Exception('The solver failed to initialize properly because: ' + reason)

    def run(self):
        """Evolve to ``tfinal`` stopping at the output times; returns the solver status dict."""
        if self.solver is None or self.solution is None:
            raise Exception('To run, a Controller must have a Solver and a Solution.')
        self.start_frame = self.solution.start_frame
        grid = self.solution.state.grid
        if len(grid.gauges) > 0:                                         # controller.py:296-297
            grid.setup_gauge_files(self.outdir)
        if not self.solver._is_set_up:
            self.solver.setup(self.solution)
            self.solver.dt = self.solver.dt_initial
        self.check_validity()
        self.solver.write_gauge_values(self.solution)                    # initial gauge values (controller.py:308-309)

        if self.output_style == 1:
            output_times = np.linspace(self.solution.t, self.tfinal, self.num_output_times + 1)
        elif self.output_style == 2:
            output_times = self.out_times
        elif self.output_style == 3:
            output_times = np.ones((self.num_output_times + 1 - self.start_frame))
        else:
            raise Exception('Invalid output style %s' % self.output_style)
        if len(output_times) == 0:
            print('No valid output times; halting.')
            return None

        frame = self.start_frame
        if self.keep_copy:
            self.frames.append(copy.deepcopy(self.solution))
        if self.output_format is not None:
            if os.path.exists(self.outdir) and self.overwrite is False:              # controller.py:331-333
                raise Exception('Refusing to overwrite existing output data. \n'
                                'Either delete/move the directory or set controller.overwrite=True.')
            self._write_p(frame)
            self.solution.write(frame, self.outdir, self.output_format, self.output_file_prefix,
                                self.write_aux_always or self.write_aux_init, self.output_options)
        self.write_F('w')
        self.log_info('Solution %s computed for time t=%f' % (frame, self.solution.t))

        status = None
        for t in output_times[1:]:
            if self.output_style < 3:
                status = self.solver.evolve_to_time(self.solution, t)
            else:
                for n in range(self.nstepout):
                    status = self.solver.evolve_to_time(self.solution)
            frame += 1
            if self.keep_copy:
                self.frames.append(copy.deepcopy(self.solution))
            if self.output_format is not None:
                self._write_p(frame)
                self.solution.write(frame, self.outdir, self.output_format, self.output_file_prefix,
                                    self.write_aux_always, self.output_options)
            self.write_F()
            self.log_info('Solution %s computed for time t=%f' % (frame, self.solution.t))
            for gfile in grid.gauge_files:                                           # controller.py:388-389
                gfile.flush()
        for gfile in grid.gauge_files:
            gfile.close()
        self.solution._start_frame = len(self.frames)
        return status

    # ---- derived quantities and functionals (controller.py:334-342, 399-409) ------------------------
    def _write_p(self, frame):
        if self.compute_p is None:
            return
        self.compute_p(self.solution.state)
        self.solution.write(frame, self.outdir_p, self.output_format, self.file_prefix_p, write_aux=False,
                            options=self.output_options, write_p=True)

    def write_F(self, mode='a'):
        if self.compute_F is None:
            return
        state = self.solution.state
        self.compute_F(state)
        F = [state.sum_F(i) for i in range(state.mF)]
        if self.is_proc_0():
            os.makedirs(self.outdir, exist_ok=True)
            with open(self.F_path, mode) as f:
                f.write(str(self.solution.t) + ' ' + ' '.join(str(j) for j in F) + '\n')

    def is_proc_0(self):
        return True

    def log_info(self, msg):
        self.logger.info(msg)
