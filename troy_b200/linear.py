"""Host mirror of ``LinearHelperCKKS::MatmulHelper`` (app/LinearHelperCKKS.cuh:104-330) over the batched C ABI.

The reference computes ``y[b][j] = sum_i x[b][i] * W[i][j]`` over a grid of CKKS ciphertexts and NTT-form plaintexts as
``B*I*J`` ``multiplyPlain`` calls plus ``B*(I-1)*J`` ``addInplace`` calls (``matmul``, :225-249). Here the grid is ONE
device buffer per operand and the whole product is one launch (``troyb_matmul_plain_ntt``). Encoding and decoding are
encoder work outside the path: this module only reproduces the block layout (which slot of which
polynomial holds which matrix entry), so any CKKS polynomial encoder can produce / read the operands.
"""
import numpy as np

from .host import CiphertextCuda, _L, _vp
from .lib import TroyInvalidArgument


def ceil_div(a, b):
    return -(-a // b)


class Cipher2d:
    """``rows x cols`` ciphertexts (Cipher2d, app/LinearHelperCKKS.cuh:34-98) as one batched ``CiphertextCuda`` whose batch
    index is ``row * cols + col``."""

    def __init__(self, rows, cols, ct):
        if ct.batch != rows * cols:
            raise TroyInvalidArgument("Not a rectangle Cipher2d.")
        self.rows, self.cols, self.ct = rows, cols, ct


class Plain2d:
    """``rows x cols`` NTT-form plaintexts (Plain2d, :17-32) as one batched ``PlaintextCuda`` (index ``row * cols + col``)."""

    def __init__(self, rows, cols, pt):
        if pt.batch != rows * cols and not (rows * cols == 1 and pt.batch == 1):
            raise TroyInvalidArgument("Not a rectangle Plain2d.")
        self.rows, self.cols, self.pt = rows, cols, pt


class MatmulHelper:
    def __init__(self, batch_size, input_dims, output_dims, slot_count):
        self.batch_size, self.input_dims, self.output_dims, self.slot_count = batch_size, input_dims, output_dims, slot_count
        self._determine_block()

    def _determine_block(self):
        """determineBlock (:113-124): the (blockHeight, blockWidth) that minimises the number of polynomials moved."""
        height, width, slots = self.input_dims, self.output_dims, self.slot_count * 2
        self.block_height = self.block_width = 0
        best = height + width + 1
        for i in range(1, height + 1):
            w = min(slots // i, width)
            if w == 0:
                break
            t = ceil_div(height, i) + ceil_div(width, w)
            if t < best:
                self.block_height, self.block_width, best = i, w, t
        # (input blocks I, output blocks J)
        self.grid = (ceil_div(height, self.block_height), ceil_div(width, self.block_width))

    # -- layout of the operands (what the polynomial encoder is fed with / what is read back)

    def weight_vectors(self, weights):
        """encodeWeights / encodeWeightSmall (:126-167): coefficient vectors [I][J][2*slot_count] of the weight blocks."""
        w = np.asarray(weights, dtype=np.float64).reshape(self.input_dims, self.output_dims)
        I, J = self.grid
        h, bw = self.block_height, self.block_width
        out = np.zeros((I, J, self.slot_count * 2))
        for bi in range(I):
            li, ui = bi * h, min((bi + 1) * h, self.input_dims)
            for bj in range(J):
                lj, uj = bj * bw, min((bj + 1) * bw, self.output_dims)
                for j in range(lj, uj):
                    for i in range(li, ui):
                        out[bi, bj, (j - lj) * h + h - (i - li) - 1] = w[i, j]
        return out

    def input_vectors(self, inputs):
        """encodeInputs (:169-194): coefficient vectors [batch][I][2*slot_count] (block i holds inputs[i*h : (i+1)*h])."""
        x = np.asarray(inputs, dtype=np.float64).reshape(self.batch_size, self.input_dims)
        I, _ = self.grid
        out = np.zeros((self.batch_size, I, self.slot_count * 2))
        for bi in range(I):
            lj, uj = bi * self.block_height, min((bi + 1) * self.block_height, self.input_dims)
            out[:, bi, : uj - lj] = x[:, lj:uj]
        return out

    def read_outputs(self, coefficient_vectors):
        """decryptOutputs (:291-311): pick y[b][j] out of decoded coefficient vectors [batch][J][2*slot_count]."""
        v = np.asarray(coefficient_vectors)
        y = np.zeros((self.batch_size, self.output_dims))
        for cid, li in enumerate(range(0, self.output_dims, self.block_width)):
            ui = min(li + self.block_width, self.output_dims)
            for j in range(li, ui):
                y[:, j] = v[:, cid, (j - li + 1) * self.block_height - 1]
        return y

    # -- the product (matmul, :225-249)
    def matmul(self, evaluator, a, weights):
        I, J = self.grid
        if a.rows != self.batch_size:
            raise TroyInvalidArgument("Input batchsize incorrect.")
        if a.cols != I or weights.rows != I or weights.cols != J:
            raise TroyInvalidArgument("Weight size incorrect.")
        x, w = a.ct, weights.pt
        if not (x.is_ntt_form and w.is_ntt_form):
            raise TroyInvalidArgument("NTT form mismatch")
        if x.parms_id != w.parms_id:
            raise TroyInvalidArgument("encrypted_ntt and plain_ntt parameter mismatch")
        scale = x.scale * w.scale
        if not evaluator._check_scale(scale, x.parms_id):
            raise TroyInvalidArgument("scale out of bounds")
        out = CiphertextCuda(x.context, x.parms_id, x.size, True, scale, x.correction_factor, batch=self.batch_size * J)
        _L().call("troyb_matmul_plain_ntt", evaluator._h(), x.parms_id, _vp(x.ptr()), x.size, x.stride, _vp(w.data.ptr),
                  _vp(out.ptr()), out.stride, self.batch_size, I, J, evaluator._s())
        return Cipher2d(self.batch_size, J, out)

    def add_plain_inplace(self, evaluator, y, x):
        """addPlainInplace (:270-289)."""
        if y.rows != x.rows or y.cols != x.cols:
            raise TroyInvalidArgument("Size incorrect.")
        evaluator.add_plain_inplace(y.ct, x.pt)

    determineBlock, addPlainInplace = _determine_block, add_plain_inplace
