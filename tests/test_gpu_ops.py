"""GPU parity tests of every C-ABI kernel against plain PyTorch fp32 ops (tolerances in tests/opcheck.py:
rel-L2 <= 1e-2 for bf16 outputs — the north-star bf16 tolerance — and tighter for fp32 outputs)."""
import pytest

import opcheck

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(opcheck.CHECKS))
def test_op(name):
    res, bad = opcheck.run(name)
    assert not bad, f"{name}: out of tolerance {bad}; all metrics {res}"
