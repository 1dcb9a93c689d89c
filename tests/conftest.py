import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as graft  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref (the real reference text compiled by oracle/Makefile)")


def _have_gpu():
    try:
        return graft.load_package().device_info()[0] > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    have = None
    for item in items:
        if "gpu" in item.keywords:
            if have is None:
                have = _have_gpu()
            if not have:
                item.add_marker(pytest.mark.skip(reason="no CUDA device"))


@pytest.fixture(scope="session")
def lab():
    """The product package (ctypes over liblabyrinth_b200.so). Fails loudly if the library is not built."""
    return graft.load_package()


@pytest.fixture(scope="session")
def O():
    """The oracle (test infrastructure only)."""
    mod = graft.load_oracle()
    mod.orc()
    return mod


@pytest.fixture(scope="session")
def ref(O):
    if not O.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return O


# libtorch's exit handlers crash now and then (SIGSEGV in c10::Dispatcher::cleanup -> free) once gloo process groups have
# been used in a process -- seen at interpreter exit of spawned workers and, 1 run in 6, of the pytest process itself, after
# every test had passed.  The session's own verdict is what the run exits with: once the summary is printed and every other
# plugin has unconfigured, leave without the handlers.
_session_status = [None]


def pytest_sessionfinish(session, exitstatus):
    _session_status[0] = int(exitstatus)


@pytest.hookimpl(trylast=True)
def pytest_unconfigure(config):
    if _session_status[0] is not None and "torch.distributed" in sys.modules:
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(_session_status[0])
