import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than a second on CPU")


@pytest.fixture(scope="session")
def golden_dir():
    return ROOT / "tests" / "golden"


OPTION_NAMES = ("render_pipe", "render_group", "render_fronts", "render_pipe_maxcap", "render_nostage", "render_mma",
                "render_mma_min", "render_mma_tmpl_min", "render_umma", "render_umma_window", "render_zero_tma", "render_umma_team", "render_rows", "render_rows_stages", "sim_lines", "sim_split", "sim_cta", "sim_stash")


@pytest.fixture
def opts():
    """Set schedule options of the native library for one test (ds_set_option); restored afterwards."""
    from diffsims_b200 import _cabi
    saved = {n: _cabi.get_option(n) for n in OPTION_NAMES}

    def set_(**kw):
        for n, v in kw.items():
            _cabi.set_option(n, v)
    yield set_
    for n, v in saved.items():
        _cabi.set_option(n, v)
