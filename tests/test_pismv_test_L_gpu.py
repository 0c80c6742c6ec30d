"""-m gpu: `pismv -test L` (test/regression/test_16.sh, isothermal SIA on a non-flat bed) time-stepped on the B200
through the C ABI: the reference's golden rows, and the oracle run tracked to rounding."""
import pytest

import cases
import gpu_util as U
import pismv_oracle as PO
from pism_b200 import icemodel

pytestmark = pytest.mark.gpu


def _device_backend(grid, cfg, inputs, max_dt):
    return icemodel.DeviceBackend(U.make_sia(grid, cfg), inputs, max_dt)


def test_pismv_test_L_golden_rows_on_device():
    for M, golden in PO.TEST_16_GOLDEN.items():
        m = PO.pismv_model("L", M, run_length_years=1000.0, backend_factory=_device_backend)
        m.run()
        o = PO.pismv_model("L", M, run_length_years=1000.0)
        o.run()
        assert m.steps == o.steps
        assert cases.rel_max(m.backend.thickness(), o.backend.thickness()) < 1e-10
        assert m.report() == golden, (M, m.report(), golden)
