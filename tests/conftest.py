import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


# The Hades tables live in the device's constant memory, one parameter set per process and device (SB200_ERR_BUSY
# otherwise), so at most one engine is alive at a time; tests that need the other round-constant rule swap it.
_live = {"rule": None, "engine": None}


def engine_for(rule):
    """The CUDA engine created with the default parameters under round-constant rule `rule`;
    fails loudly (no CPU fallback) if the library or the GPU is missing."""
    from schnorr_b200 import Engine
    if _live["engine"] is not None and _live["rule"] != rule:
        _live["engine"].close()
        _live["engine"] = None
    if _live["engine"] is None:
        _live["engine"] = Engine([0], ark=rule)
        _live["rule"] = rule
    return _live["engine"]


@pytest.fixture
def engine():
    """engine under the session's rule (environment variable SB200_ARK, default "cumsum") -- the oracle's rule"""
    import schnorr_oracle as o
    return engine_for(o.ARK_RULE)


@pytest.fixture(params=["cumsum", "plain"])
def ark(request):
    """run a test under each round-constant rule: switches the oracle (and ref_cpu / the host build through it)"""
    import schnorr_oracle as o
    prev = o.set_ark_rule(request.param)
    yield request.param
    o.set_ark_rule(prev)


@pytest.fixture
def ark_engine(ark):
    return engine_for(ark)


def pytest_sessionfinish(session, exitstatus):
    if _live["engine"] is not None:
        _live["engine"].close()
        _live["engine"] = None
