"""SURVEY.md 8a row a16 on the GPU: the product from a compact-COO operand (compactMulcsr, SimSpGEMM.cpp:247-263).
Body shared with the emulated engine: tests/test_compact.py::compact_product_check.  (The file sorts after the rest of
the GPU suite on purpose: the test was added after the round's last GPU run and has only run on the emulation.)"""
import pytest

from test_compact import compact_product_check

pytestmark = pytest.mark.gpu


def test_compact_product_on_gpu(engine):
    compact_product_check(engine)
