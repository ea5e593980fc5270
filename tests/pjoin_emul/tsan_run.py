"""Run the emulated partitioned join under ThreadSanitizer (started by
tests/test_pjoin_emul.py::test_no_data_race_under_tsan with libtsan preloaded):
a missing barrier or a missing release/acquire between work items is a data
race between the emulated threads, and would be one between CUDA threads."""
import ctypes as C
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import numpy as np                                          # noqa: E402
import test_pjoin_emul as T                                 # noqa: E402
from nbody_orbit_analysis_b200 import pjoin                 # noqa: E402
from nbody_orbit_analysis_b200.synth import SynthSim        # noqa: E402

lib = C.CDLL(sys.argv[1])
lib.pj_emul_step.argtypes = [C.POINTER(pjoin.PJoinArgs), C.c_int]
sim = SynthSim(24000, 6, 3, dtype=np.float32, catalogue_dtype=np.float32)
st = T.run_case(lib, sim, targets=[400, 400, 150], n_ctas=3, lag=1 << 11)
assert max(st['bits']) >= 4
sim = SynthSim(16000, 2, 3, dtype=np.float32, catalogue_dtype=np.float32)
st = T.run_case(lib, sim, targets=[1 << 20], n_ctas=2)
assert max(st['maxlen']) > 2 * pjoin.REC_CAP
sim = SynthSim(5000, 150, 3, dtype=np.float32, catalogue_dtype=np.float32)
st = T.run_case(lib, sim, n_ctas=3, lag=1 << 11)          # packs of tiny regions
assert max(st['packs']) >= 4
print('tsan-run-complete')
