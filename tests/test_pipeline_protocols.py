"""Random-interleaving exploration of the fused-attention kernels' mbarrier protocols (tests/pipeline_protocol_model.py).
CPU only.  The shipped protocols must run to completion with every data-hazard check silent under thousands of schedules;
the protocol of the first tensor-memory-P version (one P-full barrier per half) must be caught."""
import pytest

from pipeline_protocol_model import Hazard, variant1, variant2


@pytest.mark.parametrize("nb", [1, 2, 3, 4, 8])
def test_variant1_protocol_has_no_deadlock_or_hazard(nb):
    for seed in range(300):
        variant1(nb, seed).run()


@pytest.mark.parametrize("kst,vst", [(2, 1), (2, 2), (4, 3)])
def test_variant1_protocol_other_ring_depths(kst, vst):
    for seed in range(150):
        variant1(5, seed, kst=kst, vst=vst).run()


def test_model_catches_the_single_pfull_barrier_race():
    """The version that hung under CUDA-graph replay: softmax of block j+1 can complete a second phase of the per-half
    P-full barrier before the MMA warp has looked at the first; a parity wait cannot see two flips."""
    caught = 0
    for seed in range(400):
        try:
            variant1(6, seed, pfull_per_buffer=False).run()
        except Hazard:
            caught += 1
    assert caught > 0, "the model no longer reproduces the known protocol bug"


@pytest.mark.parametrize("nb,items", [(1, 1), (1, 3), (2, 2), (3, 3), (5, 2), (8, 2)])
def test_variant2_protocol_has_no_deadlock_or_hazard(nb, items):
    for seed in range(250):
        variant2(nb, items, seed).run()


@pytest.mark.parametrize("kst,vst", [(2, 1), (2, 2), (4, 3)])
def test_variant2_protocol_other_ring_depths(kst, vst):
    for seed in range(100):
        variant2(4, 3, seed, kst=kst, vst=vst).run()


def _mutant(old, new, which, args, seeds=120):
    """Runs the model with one line of its source replaced; returns how many schedules raised a Hazard."""
    import inspect

    import pipeline_protocol_model as m
    src = inspect.getsource(m)
    assert old in src
    ns = {}
    exec(compile(src.replace(old, new, 1), "attn_protocol_mutant", "exec"), ns)
    caught = 0
    for seed in range(seeds):
        try:
            ns[which](*args, seed).run()
        except ns["Hazard"]:
            caught += 1
    return caught


@pytest.mark.parametrize("label,old,new,which,args", [
    ("v2: S product not waiting for the previous scores/P to be consumed",
     '            if sc[x] > 0:\n                yield M.wait(f"SDONE{x}", (sc[x] - 1) & 1)\n',
     '            if False:\n                yield None\n', "variant2", (3, 3)),
    ("v2: P V not waiting for P", '                    yield M.wait(f"SDONE{x}", (sc[x] - 1) & 1)\n                    if j == 0:',
     '                    if j == 0:', "variant2", (3, 3)),
    ("v2: Q replaced without waiting for the last score products", '            yield M.wait("QEMPTY", (wi & 1) ^ 1)\n',
     '            pass\n', "variant2", (3, 3)),
    ("v2: S-done barrier counting 3 of the 4 warps", 'M.bar(f"SFULL{x}", 1); M.bar(f"SDONE{x}", 4)',
     'M.bar(f"SFULL{x}", 1); M.bar(f"SDONE{x}", 3)', "variant2", (3, 3)),
    ("v1: S-empty barrier without the MMA warp's arrival", 'M.bar(f"SFULL{s}", 1); M.bar(f"SEMPTY{s}", 9)',
     'M.bar(f"SFULL{s}", 1); M.bar(f"SEMPTY{s}", 8)', "variant1", (5,)),
])
def test_model_is_sensitive_to_protocol_mutations(label, old, new, which, args):
    assert _mutant(old, new, which, args) > 0, label


@pytest.mark.parametrize("n_tiles,num_kb,stages", [(1, 1, 3), (3, 16, 3), (4, 9, 6), (2, 20, 3), (5, 3, 2)])
def test_gemm_pipeline_protocol(n_tiles, num_kb, stages):
    """gemm_tc_kernel: TMA ring, K-chunk promotion through the two TMEM accumulators, 8 epilogue warps."""
    from pipeline_protocol_model import gemm
    for seed in range(150):
        gemm(n_tiles, num_kb, seed, stages=stages).run()
