"""GPU parity of whole training-step graphs vs the CPU oracle (identical weights, batch and noise)."""
import pytest

from tests import parity as P

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("H,C,L,B", [(32, 3, 16, 8), (32, 3, 200, 8), (64, 3, 16, 4), (32, 3, 16, 64)])
def test_iwgan_step_matches_oracle(H, C, L, B):
    res = P.iwgan_step_parity(H=H, C=C, L=L, B=B, verbose=True)
    assert res["ok"], res
