"""`-m "not gpu"`: the N > 1 path on CPU: world_size-2 `gloo` process groups drive the SAME host code as on GPUs
(troy_b200/distributed.py), with the library's kernels executed by tests/emul (a debugging aid, never the product) so that
the sharded result can be compared word for word with the reference."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_everything():
    from troy_b200.distributed import shard_range
    for n in (0, 1, 7, 8, 40, 128):
        for world in (1, 2, 3, 4, 8):
            parts = [shard_range(n, world, r) for r in range(world)]
            assert [i for p in parts for i in p] == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, scheme, n, bits, mode, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", world_size=world, rank=rank)
    try:
        import troy_b200 as tb
        from troy_b200 import lib as tlib
        from troy_b200.distributed import DigitShardedKeySwitch, shard_range
        from oracle import ref
        from common import make_pair, random_ct_words
        tlib._library = tlib.Library(os.path.join(ROOT, "tests", "emul", "libtroy_b200_emul.so"))
        tb.KernelProvider.initialize(0)
        rc, ctx = make_pair(ref, tb, scheme, n, bits)
        rc.keygen()                      # seeded: identical keys in every rank
        rlk = rc.relin_keys()
        key = rlk.words(0)
        L = ctx.level_limbs(0)
        mine = shard_range(L, world, rank)
        ks = DigitShardedKeySwitch(ctx, key[mine.start:mine.stop] if len(mine) else None, mode=mode)
        ntt = scheme == ref.CKKS
        w = random_ct_words(rc, 0, 3, n, np.random.default_rng(5), batch=2)   # same inputs in every rank
        ct = tb.CiphertextCuda.from_host(ctx, w, 0, ntt, 2.0 ** 30 if ntt else 1.0)
        ks.relinearize_inplace(ct)
        got = ct.to_host()
        for b in range(2):
            r = rc.ciphertext(w[b], 0, ntt, 2.0 ** 30 if ntt else 1.0)
            rc.eval(ref.OP_RELIN, r, keys=rlk)
            assert np.array_equal(got[b], r.words()), f"rank {rank}: sharded relinearize differs from the reference"
        out.put((rank, "ok"))
    except Exception as e:  # report instead of hanging the peer in a collective
        out.put((rank, f"{type(e).__name__}: {e}"))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["reduce_scatter", "all_reduce"])
@pytest.mark.parametrize("scheme,n,bits", [(2, 4096, [50, 40, 40, 40, 50]), (1, 4096, [40, 36, 36, 40])])
def test_digit_sharded_key_switch_world_size_2(ref, scheme, n, bits, mode):
    import torch.multiprocessing as mp
    import __graft_entry__ as g
    g.build_emul()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, scheme, n, bits, mode, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    results = dict(out.get(timeout=5) for _ in procs)
    assert results == {0: "ok", 1: "ok"}, results
    assert all(p.exitcode == 0 for p in procs)
