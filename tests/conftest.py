import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


# the suites pin the precision of every trainer explicitly (set_precision); a trainer that does not is the reference-exact
# fp32 mode here, whatever the user's environment says ('auto' = bf16 where implemented is the product default)
os.environ["MCPC_PRECISION"] = "fp32"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box via gpurun)")
