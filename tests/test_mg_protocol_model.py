"""The ordering rules of the multi-GPU layer's peer-memory transports (csrc/mg.cu), executed as a model under an adversarial scheduler
(tests/mg_protocol_model.py): no deadlock, no unordered conflicting accesses to staging / plane buffers or to the caller's B, every split
and GEMM sees the data of its own product -- at rank counts (8) and in combinations (root-free ranges, changing roots, distributed B) that
the GPU self-test (tools/mg_selftest.py, 2 and 4 ranks) could not all visit.  CPU only."""
import pytest

import mg_protocol_model as M


@pytest.mark.parametrize("transport", ["raw", "push", "planes"])
@pytest.mark.parametrize("nr,root,root_free_min", [(2, 0, 6), (2, 1, 2), (3, 2, 6), (4, 0, 6), (4, 3, 4), (8, 0, 6), (8, 5, 6), (8, 0, 0), (4, -1, 6), (8, -1, 6)])
@pytest.mark.parametrize("caller", ["rewrite", "const", "inorder"])
def test_protocol_is_race_free_and_deadlock_free(transport, nr, root, root_free_min, caller):
    for seed in range(3):
        steps = M.simulate(nr, transport, root=root, products=6, caller=caller, root_free_min=root_free_min, seed=seed)
        assert steps > 0


@pytest.mark.parametrize("transport", ["raw", "push", "planes"])
def test_any_rank_may_be_the_root_of_a_later_product(transport):
    """staging buffers are released to EVERY rank (F_SPLIT_DONE goes to all), so the root may change from product to product"""
    for seed in range(4):
        M.simulate(4, transport, products=8, caller="rewrite", roots=[0, 3, 1, 1, 2, 0, 3, 2], seed=seed)
        M.simulate(8, transport, products=8, caller="rewrite", roots=[0, 7, 1, 6, 2, 5, 3, 4], seed=seed)


@pytest.mark.parametrize("transport", ["raw", "push", "planes"])
def test_empty_ranges_and_few_columns(transport):
    """fewer 256-column blocks than ranks: some owners have nothing to send, their flags must not be waited for in vain"""
    for blocks in (1, 3, 5):
        for seed in range(3):
            M.simulate(8, transport, products=5, caller="rewrite", blocks=blocks, seed=seed)
            M.simulate(4, transport, root=-1, products=5, caller="rewrite", blocks=blocks, seed=seed)


@pytest.mark.parametrize("transport,broken", [("raw", "no_buffer_reuse_wait"), ("push", "no_buffer_reuse_wait"), ("planes", "no_buffer_reuse_wait"),
                                              ("raw", "no_caller_fence"), ("raw", "no_forward_join")])
def test_the_model_has_teeth(transport, broken):
    """removing one ordering rule (the epoch e-3 buffer-reuse waits; the fence between the copies that read B and the caller's next
    write; the join of the forwards before the staging buffer is declared consumed) must be detected in at least one schedule.
    (On the plane transports the caller fence is implied by data flow whenever the root multiplies every range -- its GEMMs need planes
    that only exist once the scatter has been consumed -- so only the raw transport, whose root splits straight from B, is probed.)"""
    caught = 0
    for seed in range(12):
        try:
            M.simulate(4, transport, products=8, caller="rewrite", seed=seed, broken=broken)
        except M.ProtocolError:
            caught += 1
    assert caught > 0, f"{transport}/{broken}: the model did not notice the missing rule"


def test_ranges_match_the_library():
    """the model's range helper == gffm_mg_owner_ranges / gffm_mg_owner_ranges_root_free (in 256-column blocks)"""
    import gffm_b200 as g
    for n_blocks, nr, root in [(64, 8, 0), (64, 8, 7), (9, 4, 2), (3, 8, 1), (17, 6, 5)]:
        off = g.multigpu.owner_ranges_root_free(256 * n_blocks, nr, root)
        assert [(b - a) // 256 for a, b in zip(off, off[1:])] == M.ranges(nr, root, True, n_blocks)
        off = g.multigpu.owner_ranges(256 * n_blocks, nr)
        assert [(b - a) // 256 for a, b in zip(off, off[1:])] == M.ranges(nr, 0, False, n_blocks)
