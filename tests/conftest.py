import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


# The GPU tests run up to six y-slabs (engine handles, two streams each) on ONE device, with device-side handshake kernels that spin
# until another slab's kernel has run.  With the default 8 hardware queues two such streams can share a queue, and the spinning
# kernel then blocks the one it waits for until the 10 s handshake timeout.  On an 8-GPU box every slab has its own device.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")
