"""Run under torchrun on N GPUs of one node (tests/test_multi_gpu.py launches it; also usable by hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py

Checks, against the reference CPU library on every rank: (1) a digit-sharded relinearization (NCCL reduce-scatter of the
row-major partial key inner products, broadcast of the special row, all-gather of the finished limbs; and the all-reduce
variant) is word-for-word the reference's result; (2) independent-ciphertext sharding: every rank's
slice of a batch equals the reference."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import troy_b200 as tb
    from troy_b200.distributed import DigitShardedKeySwitch, shard_range
    from oracle import ref
    from common import make_pair, random_ct_words
    tb.KernelProvider.initialize(local)
    stream = torch.cuda.current_stream().cuda_stream
    for scheme, n, bits in [(ref.CKKS, 16384, [60, 40, 40, 40, 40, 40, 60]), (ref.BGV, 8192, [50, 40, 40, 50])]:
        rc, ctx = make_pair(ref, tb, scheme, n, bits)
        rc.keygen()
        rlk = rc.relin_keys()
        L = ctx.level_limbs(0)
        mine = shard_range(L, world, rank)
        ntt = scheme == ref.CKKS
        sc = 2.0 ** 30 if ntt else 1.0
        batch = 4
        w = random_ct_words(rc, 0, 3, n, np.random.default_rng(11), batch=batch)
        want = []
        for b in range(batch):
            r = rc.ciphertext(w[b], 0, ntt, sc)
            rc.eval(ref.OP_RELIN, r, keys=rlk)
            want.append(r.words())
        want = np.stack(want)
        got = None
        for mode in ("reduce_scatter", "all_reduce"):
            ks = DigitShardedKeySwitch(ctx, rlk.words(0)[mine.start:mine.stop] if len(mine) else None, stream=stream, mode=mode)
            ct = tb.CiphertextCuda.from_host(ctx, w, 0, ntt, sc)
            ks.relinearize_inplace(ct)
            torch.cuda.synchronize()
            got = ct.to_host()
            bad = [(int(b), int(k), int(j)) for b, k, j in zip(*np.nonzero((got != want).any(axis=3)))]
            assert not bad, f"rank {rank}: digit-sharded relinearize ({mode}) differs from the reference at (batch, poly, limb) {bad[:12]} (scheme {scheme})"
        # (2) each rank also checks only ITS ciphertexts of the batch through the unsharded path
        ev = tb.EvaluatorCuda(ctx, stream=stream)
        full = tb.RelinKeysCuda.from_host(ctx, {0: rlk.words(0)})
        for b in shard_range(batch, world, rank):
            r = rc.ciphertext(w[b], 0, ntt, sc)
            rc.eval(ref.OP_RELIN, r, keys=rlk)
            assert np.array_equal(got[b], r.words()), f"rank {rank}: digit-sharded relinearize differs (scheme {scheme})"
            mine_ct = tb.CiphertextCuda.from_host(ctx, w[b], 0, ntt, sc)
            ev.relinearize_inplace(mine_ct, full)
            ev.synchronize()
            assert np.array_equal(mine_ct.to_host()[0], r.words()), f"rank {rank}: ciphertext shard differs (scheme {scheme})"
    ok = torch.ones(1, device="cuda")
    dist.all_reduce(ok)
    if rank == 0:
        print(f"multi_gpu_check ok on {int(ok.item())} ranks", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
