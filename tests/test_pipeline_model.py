"""The GEMM's synchronisation protocol (vit.rs_b200/csrc/gemm_tc.cu) under a random scheduler: tests/pipeline_model.py.

The regimes are those of the training step: K = 192 (3 k-blocks per tile: ViT-Ti/16, ViT-B/8's patch embedding) to K = 3072 (48),
thousands of tiles per launch scaled down to tens per cluster, split-K weight gradients with about one unit per cluster,
epilogues much slower than the main loop (K = 192 with a fused epilogue) and the reverse, late TMA / st.async / remote-arrive
completions, CTA pairs and single CTAs, the dynamic (atomic counter) and the static tile order, two launches in a row (the
counter must be re-armed)."""
import pytest

from tests import pipeline_model as pm

REGIMES = {
    "k192-many-tiles": dict(total_units=40, kb=3, clusters=3),
    "k192-slow-epilogue": dict(total_units=40, kb=3, clusters=3, slow={"epilogue": 0.9}),
    "k192-slow-epilogue-deep": dict(total_units=60, kb=3, clusters=2, slow={"epilogue": 0.95, "event": 0.5}, epi_cost=6),
    "k768-late-completions": dict(total_units=30, kb=12, clusters=2, slow={"event": 0.8}),
    "k3072": dict(total_units=8, kb=48, clusters=2),
    "k64-one-block": dict(total_units=30, kb=1, clusters=3),
    "fewer-units-than-clusters": dict(total_units=2, kb=5, clusters=3),
    "one-unit-per-cluster": dict(total_units=3, kb=7, clusters=3),
    "slow-producer": dict(total_units=50, kb=3, clusters=2, slow={"producer": 0.9}),
    "slow-tensor-pipe": dict(total_units=50, kb=3, clusters=2, slow={"pipe": 0.9, "epilogue": 0.5}),
    "single-cta-4-stages": dict(total_units=25, kb=3, clusters=2, CG=1, STAGES=4),
    "single-cta-6-stages": dict(total_units=25, kb=2, clusters=2, CG=1),
    "static-order": dict(total_units=30, kb=3, clusters=3, dynamic=False),
    "no-units-at-all": dict(total_units=0, kb=3, clusters=2),
}


@pytest.mark.parametrize("name", sorted(REGIMES))
def test_protocol_has_no_deadlock_aliasing_or_hazard(name):
    for seed in range(8):
        assert pm.simulate(1000 * seed + len(name), **REGIMES[name])


@pytest.mark.parametrize("fault,units,kb", [("tempty_short", 30, 3), ("no_sempty_wait", 60, 1), ("empty_parity", 10, 3)])
def test_the_checker_notices_a_broken_protocol(fault, units, kb):
    """Each fault is a plausible slip of the kernel source: an arrival count one short (the MMA issuer may overwrite an
    accumulator a warp is still reading), a producer that does not wait for the tile-queue slot, a wrong initial parity."""
    caught = 0
    for seed in range(6):
        try:
            pm.simulate(seed, total_units=units, kb=kb, clusters=2, slow={"epilogue": 0.9}, fault=fault)
        except pm.ProtocolError:
            caught += 1
    assert caught >= 5, caught
