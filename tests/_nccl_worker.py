"""2-GPU worker for tests/test_gpu_multi.py: native NCCL gather through the C-ABI vs single-GPU results."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from genomics_general_b200 import multigpu, synth  # noqa: E402
from genomics_general_b200.engine import Engine, PinnedArray  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group(backend="nccl", device_id=dev)
    eng = Engine(local)
    id_t = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        id_t.copy_(torch.frombuffer(bytearray(eng.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(id_t, 0)
    eng.nccl_init(world, rank, bytes(id_t.cpu().numpy().tobytes()))
    for miss in (0.0, 0.03):
        spec = synth.SynthSpec(4, 10, miss=miss, seed=3)
        S = 60000
        g = synth.synth_genotypes(spec, 0, S)
        pos = synth.synth_positions(S)
        lo = np.arange(0, S, 1700, dtype=np.int64)
        hi = np.minimum(lo + 1700, S)
        shards = multigpu.shard_windows(lo, hi, world)
        b, e = shards[rank]
        s0, s1 = multigpu.shard_site_range(lo, hi, b, e)
        eng.upload(g[s0:s1], pos[s0:s1])                       # each rank holds only its shard of the sites
        eng.set_pops(spec.hap_pop(), 4)
        eng.set_windows(lo[b:e] - s0, hi[b:e] - s0)
        counts = [x[1] - x[0] for x in shards]
        w_max = max(counts)
        table = PinnedArray((world * w_max, eng.popgen_record_width()), np.float64)
        nk2 = eng.popgen_allgather(w_max, table.array, 100, 0.01)
        rows = np.concatenate([table.array[r * w_max: r * w_max + counts[r]] for r in range(world)], axis=0)
        got = multigpu.unpack_device_records(rows, 4)
        # single-GPU truth on this rank
        eng.upload(g, pos)
        eng.set_pops(spec.hap_pop(), 4)
        eng.set_windows(lo, hi)
        ref = eng.popgen(100, 0.01)
        assert np.array_equal(got["sites"], ref["sites"]) and np.array_equal(got["pos_sum"], ref["pos_sum"])
        assert np.array_equal(got["path"], ref["path"])
        assert (nk2 > 0) == (miss > 0)
        for k in ("pi", "dxy", "fst"):
            assert np.allclose(got[k], ref[k], rtol=1e-12, atol=0, equal_nan=True), k
        table.close()
        # ABBA-BABA and fourPop, window-sharded the same way (config 3)
        eng.upload(g[s0:s1], pos[s0:s1])
        eng.set_pops(spec.hap_pop(), 4)
        eng.set_windows(lo[b:e] - s0, hi[b:e] - s0)
        ta = PinnedArray((world * w_max, 8), np.float64)
        eng.abbababa_allgather(0, 1, 2, 3, 0.5, w_max, ta.array)
        ga = multigpu.unpack_abba_records(multigpu.gathered_rows(ta.array, counts, w_max))
        tf = PinnedArray((world * w_max, 17), np.float64)
        eng.fourpop_allgather(0, 1, 2, 3, 0.5, w_max, tf.array, polarize=True)
        gf = multigpu.unpack_fourpop_records(multigpu.gathered_rows(tf.array, counts, w_max))
        eng.upload(g, pos)
        eng.set_pops(spec.hap_pop(), 4)
        eng.set_windows(lo, hi)
        ra = eng.abbababa(0, 1, 2, 3, 0.5)
        rf = eng.fourpop(0, 1, 2, 3, 0.5, polarize=True)
        for k in ("sites", "pos_sum"):
            assert np.array_equal(ga[k], ra[k]) and np.array_equal(gf[k], rf[k]), k
        # fp64 sums: a shard tiles its sites differently, so the summation order differs from the single-GPU run
        for k in ("ABBA", "BABA", "D", "fd", "fdM", "sitesUsed"):
            assert np.allclose(ga[k], ra[k], rtol=1e-11, atol=1e-13, equal_nan=True), k
        for k in multigpu.FOURPOP_KEYS + ("sitesUsed",):
            assert np.allclose(gf[k], rf[k], rtol=1e-10, atol=1e-13, equal_nan=True), k
        ta.close()
        tf.close()
    # ---- distMat --windType cat: the single window is sharded along the SITE axis; one ncclAllReduce of the
    # integer pair matrices ----
    spec = synth.SynthSpec(3, 6, miss=0.04, seed=8)
    S = 150001
    g = synth.synth_genotypes(spec, 0, S)
    hap_ind = (np.arange(g.shape[1]) // 2).astype(np.int32)
    n_ind = g.shape[1] // 2
    s0, s1 = rank * S // world, (rank + 1) * S // world
    eng.upload(g[s0:s1], None)
    got, tot = eng.pairdist_cat(hap_ind, n_ind, False)
    assert tot == S
    eng.nccl_finalize()                      # without a communicator the same call covers only the local sites
    eng.upload(g, None)
    eng.set_windows([0], [S])
    ref = eng.pairdist(hap_ind, n_ind, False)["dist"][0]
    assert np.array_equal(got, ref, equal_nan=True)          # integer sums, one division: bit-identical
    dist.barrier()
    eng.close()
    dist.destroy_process_group()
    if rank == 0:
        print("NCCL_GATHER_OK")


if __name__ == "__main__":
    main()
