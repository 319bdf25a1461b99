"""The persistent attention backward kernel's protocol (attn_bwd_persist_kernel) under a random scheduler, for every sequence
length it accepts: tests/attn_bwd_model.py.  The GPU suite runs the kernel at a handful of lengths (T = 197, 256, ...); the parity
formulas of its barriers depend on T through the tile counts (NT = 1 or 2 key tiles, NSUB = 1..4 query sub-tiles, odd key tiles
visiting their sub-tiles in the order 0, 1, 3, 2), so every T is modelled here."""
import pytest

from tests import attn_bwd_model as am


def test_every_sequence_length():
    for T in range(1, 257):
        for nheads, seed in ((1, 0), (3, 1), (4, 2)):
            assert am.simulate(1000 * T + seed, T, nheads)


@pytest.mark.parametrize("T", [1, 64, 65, 128, 129, 192, 193, 197, 256])
@pytest.mark.parametrize("slow", [{"group": 0.9}, {"readout": 0.95}, {"loader": 0.95, "event": 0.8}, {"issuer": 0.9}, {"pipe": 0.9},
                                  {"statistics": 0.97}, {"group": 0.5, "readout": 0.5, "event": 0.9}],
                         ids=["slow-groups", "slow-readout", "slow-loads", "slow-issuer", "slow-pipe", "slow-statistics", "mixed"])
def test_tile_boundaries_with_slow_roles(T, slow):
    for seed in range(3):
        assert am.simulate(seed * 31 + T, T, 6, slow=slow, work=3)


@pytest.mark.parametrize("fault", ["acc_free_short", "no_q0_wait"])
def test_the_checker_notices_a_broken_protocol(fault):
    """An arrival count one short (dV / dK restarted before the read-out group has them), a loader that refills Q_0 / dO_0 without
    waiting for the last MMAs that read them."""
    caught = 0
    for seed in range(8):
        try:
            am.simulate(seed, 197, 5, slow={"readout": 0.9, "pipe": 0.8}, fault=fault)
        except am.ProtocolError:
            caught += 1
    assert caught >= 6, caught
