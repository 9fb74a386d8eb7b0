import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "dmft-lanc-ed_b200"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "multigpu: needs >= 2 GPUs (launched under torchrun by tests/run_multigpu.sh)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The oracle is compiled on demand (gcc only); libedgpu.so must have been built by
    __graft_entry__.build() -- tests fail loudly if it is missing."""
    import oracle
    oracle.build()
    yield


def make_oracle(name, **over):
    import oracle as O
    from edgpu import configs
    cfg = configs.config(name)
    cfg.update(over)
    return cfg, O.Oracle(**configs.solver_kwargs(cfg))
