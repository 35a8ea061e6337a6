"""Multi-GPU layer of the RNS polynomial path. One process per GPU, torch.distributed for the plumbing.

Two modes, matching how the path shards:

* **independent ciphertexts** (the default and the graded mode): every evaluator call touches only its own ciphertexts plus
  read-only tables and keys, so a batch is split across ranks with :func:`shard_range`, tables and keys are replicated,
  and there is NO collective on the data path (``bench.py --gpus N`` does exactly this);
* **one key switch sharded by digit** (large N, many limbs: BASELINE.json config 5, where one key is 1.6 GiB): rank g holds
  the key digits ``shard_range(L, world, g)`` and runs the mod-up transforms and the key inner product for those digits only.
  The partial products are kept ROW-major (``[L+1 output moduli][batch][2][N]``, ``troyb_switch_key_partial_rows``) and summed
  with ONE reduce-scatter (NCCL over NVLink / NVSwitch; exact: at most 8 canonical residues below 2^60 fit 64 bits), so that
  every rank receives, reduces and mod-downs only the rows (output moduli) it owns; the rounded special-prime row - 2 N words
  per ciphertext, which every data limb's mod-down needs - is broadcast by its owner, and the finished limbs are all-gathered
  into every rank's replica of the ciphertext. Everything is enqueued on one stream: no host synchronisation between the
  kernels and the collectives. (``mode="all_reduce"`` keeps the simpler variant: one all-reduce of the whole product, every
  rank finishing the whole mod-down.) The reference has no multi-GPU code at all.
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
    None when the slice is empty; ``key_device_ptr`` instead when the slice is already on the device). ``group=None`` is the
    default group. ``stream`` is the CUDA stream the kernels run on: torch must be on the same stream when the default stream
    is used (None or 0); any other stream is made torch's current stream for the duration of a key switch."""

    def __init__(self, context, key_words, level=0, group=None, stream=None, mode="reduce_scatter", key_device_ptr=None):
        import torch.distributed as dist
        self.dist = dist
        self.context, self.level, self.group, self.stream = context, level, group, stream
        if mode not in ("reduce_scatter", "all_reduce"):
            raise TroyInvalidArgument("mode must be 'reduce_scatter' or 'all_reduce'")
        self.mode = mode
        self.bytes_sent = 0  # NVLink bytes this rank sent in the last key switch (algorithmic: ring reduce-scatter / all-gather)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise TroyInvalidArgument("at most 8 partial products can be summed in 64 bits")
        L = context.level_limbs(level)
        r = shard_range(L, self.world, self.rank)
        self.digit_begin, self.digit_end = r.start, r.stop
        key_limbs = context.level_limbs(SEALContextCuda.KEY_LEVEL)
        self.digit_words = 2 * key_limbs * context.n
        self.key = None
        if key_device_ptr is not None:
            # this rank's slice already lives on the device (the caller keeps it alive)
            from .host import ExternalArray
            self.key = ExternalArray(int(key_device_ptr), len(r) * self.digit_words)
        elif len(r):
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
        if self.world > 1 and self.mode == "reduce_scatter":
            return self._switch_key_rows(ct, target_ptr, target_stride)
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

    def _on_stream(self):
        """Context manager that makes the evaluator stream torch's current stream, so that torch's fills, copies and
        collectives are ordered with the library's kernels without a host synchronisation."""
        import contextlib
        import torch
        if not self.stream or not torch.cuda.is_available():
            return contextlib.nullcontext()  # the default stream: torch must be on it as well
        return torch.cuda.stream(torch.cuda.ExternalStream(int(self.stream)))

    def _ct_view(self, ct):
        """The ciphertext words as a torch tensor [batch][capacity][L][N] over the same memory (no copy)."""
        import numpy as np
        import torch
        shape = (ct.batch, ct.capacity, ct.limbs, self.context.n)
        if torch.cuda.is_available():
            class _Cai:  # __cuda_array_interface__ view of memory the library owns
                pass
            v = _Cai()
            v.__cuda_array_interface__ = {"shape": shape, "typestr": "<i8", "data": (int(ct.ptr()), False), "version": 2}
            return torch.as_tensor(v, device=torch.device("cuda", torch.cuda.current_device()))
        words = int(np.prod(shape))
        buf = (C.c_int64 * words).from_address(int(ct.ptr()))
        return torch.from_numpy(np.frombuffer(buf, dtype=np.int64).reshape(shape))

    def _switch_key_rows(self, ct, target_ptr, target_stride):
        """The reduce-scatter variant (see the module docstring)."""
        import torch
        ctx, dist, lib = self.context, self.dist, _lib.library()
        L, n, B, W = ctx.level_limbs(self.level), ctx.n, ct.batch, self.world
        per = -(-(L + 1) // W)                     # rows per rank (the last ranks' blocks may be short or empty)
        r0 = min(self.rank * per, L + 1)
        r1 = min(r0 + per, L + 1)
        row_words = B * 2 * n
        dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        nccl = dist.get_backend(self.group) == "nccl"
        with self._on_stream():
            s = _vp(self.stream)
            rows = torch.zeros(per * W, B, 2, n, dtype=torch.int64, device=dev)
            if self.digit_end > self.digit_begin:
                lib.call("troyb_switch_key_partial_rows", _vp(ctx.h), self.level, _vp(target_ptr), target_stride,
                         _vp(self._key_base()), self.digit_begin, self.digit_end, _vp(rows.data_ptr()), B, s)
            if nccl:
                mine = torch.empty(per, B, 2, n, dtype=torch.int64, device=dev)
                dist.reduce_scatter_tensor(mine, rows, op=dist.ReduceOp.SUM, group=self.group)
            else:  # gloo has no reduce-scatter: the same sums through an all-reduce (host-side tests)
                dist.all_reduce(rows, op=dist.ReduceOp.SUM, group=self.group)
                mine = rows[self.rank * per:(self.rank + 1) * per].contiguous()
            # `mine` holds rows [r0, r0 + per); the library indexes rows from row 0
            virt = mine.data_ptr() - 8 * r0 * row_words
            if r1 > r0:
                lib.call("troyb_switch_key_reduce_rows", _vp(ctx.h), self.level, _vp(virt), r0, r1, B, s)
            sp = torch.empty(B, 2, n, dtype=torch.int64, device=dev)
            owner = L // per
            if self.rank == owner:
                lib.call("troyb_switch_key_special_row", _vp(ctx.h), self.level, _vp(virt), _vp(sp.data_ptr()), B, s)
            dist.broadcast(sp, src=dist.get_global_rank(self.group, owner) if self.group is not None else owner, group=self.group)
            j1 = min(r1, L)
            if j1 > r0:
                lib.call("troyb_switch_key_moddown_rows", _vp(ctx.h), self.level, _vp(ct.ptr()), ct.stride, _vp(virt),
                         _vp(sp.data_ptr()), r0, j1, B, s)
            # every rank's replica of the ciphertext receives the limbs the others finished
            view = self._ct_view(ct)
            send = torch.zeros(per, B, 2, n, dtype=torch.int64, device=dev)
            if j1 > r0:
                send[:j1 - r0] = view[:, :2, r0:j1].permute(2, 0, 1, 3)
            full = torch.empty(per * W, B, 2, n, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(full, send, group=self.group)
            view[:, :2] = full[:L].permute(1, 2, 0, 3)
            if dev.type == "cuda":
                # the temporaries are torch tensors: the caching allocator keeps them for this stream
                for t in (rows, mine, sp, send, full):
                    t.record_stream(torch.cuda.current_stream())
        self.bytes_sent = 8 * ((W - 1) * per * row_words * 2 + (B * 2 * n if self.rank == owner else 0))

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
