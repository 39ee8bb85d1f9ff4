"""The driver's smoke() entry point is part of the GPU suite, so that it cannot rot unnoticed."""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


@pytest.mark.gpu
def test_graft_entry_smoke(capsys):
    import __graft_entry__ as entry
    entry.smoke()
    assert "smoke ok" in capsys.readouterr().out
