"""LayerNorm — drop-in for the reference's nn/normalization.py."""
from __future__ import annotations

from .. import ops
from .. import runtime as rt
from .core import Layer, weight_init_fn


class LayerNorm(Layer):
    def __init__(self, d_model: int, eps: float = 1e-5, weight_init: str = "scaled_normal",
                 bias_init: str = "zeros", **kwargs):
        super().__init__()
        scale = kwargs.pop("scale", 0.2)  # nn/normalization.py:17 — weight is NOT ones
        self.add_settings({"d_model": d_model, "eps": eps})
        self.add_parameter("weight", weight_init_fn(weight_init, **kwargs, scale=scale)((d_model,)))
        self.add_parameter("bias", weight_init_fn(bias_init, **kwargs)((d_model,)))

    def forward(self, inputs, residual=None):
        """gamma*(x-mean)/sqrt(var+eps)+beta, optionally + residual in the same pass
        (TransformerBlock's `norm(sub(x)) + x`, transformer_block.py:20-21)."""
        pinfo = getattr(inputs, "_npt_peer", None)
        if pinfo is not None:
            # tensor parallelism over 2 ranks: `inputs` is this rank's partial, the other rank's is read from
            # peer memory inside the kernel (no all-reduce kernel; distributed/peer.py)
            px, peer_ptr = pinfo
            px.barrier()
            rows = getattr(inputs, "_npt_rows", None)
            if rows is not None:
                # row-sharded: normalise the rows this rank owns, push the bf16 result into the other rank's copy,
                # meet again — both ranks then hold every row of the tensor the next GEMM reads
                outputs, mean, var, inputs = ops.layernorm_fwd_rows(inputs, peer_ptr, self.weight.data, self.bias.data,
                                                                    self.eps, residual, rows, px)
                px.barrier()
                self.ctx |= {"inputs": inputs, "mean": mean, "var": var, "rows": rows, "px": px}
                self.ran_row_sharded = True
                return outputs
            outputs, mean, var, inputs = ops.layernorm_fwd(inputs, self.weight.data, self.bias.data, self.eps,
                                                           residual=residual, want_bf16=rt.is_bf16(), x_peer=peer_ptr)
        else:
            outputs, mean, var = ops.layernorm_fwd(inputs, self.weight.data, self.bias.data, self.eps,
                                                   residual=residual, want_bf16=rt.is_bf16())
        self.ctx |= {"inputs": inputs, "mean": mean, "var": var}
        return outputs

    def backward(self, d_out, gemm_only: bool = False):
        """`gemm_only`: the caller feeds d_in straight into Linear.backward, so bf16 mode may return
        it as a bf16 tensor alone, and that layer's bias gradient (the column sums of d_in) is computed
        here, in the pass that produces d_in."""
        inputs, mean, var = self.ctx.pop("inputs"), self.ctx.pop("mean"), self.ctx.pop("var")
        rows, px = self.ctx.pop("rows", None), self.ctx.pop("px", None)
        pinfo = getattr(d_out, "_npt_peer", None)
        peer_ptr = 0
        if rows is not None:
            # the forward pass was row-sharded (only the own rows of `inputs` / mean / var exist): so is this one
            assert pinfo is not None and getattr(d_out, "_npt_rows", None) == rows and gemm_only, \
                "row-sharded LayerNorm.backward needs the split-stored partial dX of the layer above"
            px.barrier()
            d_in, finish = ops.layernorm_bwd_rows(inputs, self.weight.data, mean, var, d_out, pinfo[1], self.eps, rows, px,
                                                  want_colsum=True)
            px.barrier()
            d_weight, d_bias, colsum = finish()
            d_in._npt_colsum = colsum
            self.update_parameter("weight", gradient=d_weight)
            self.update_parameter("bias", gradient=d_bias)
            return d_in
        if pinfo is not None:  # d_out is this rank's partial dX; the other rank's is added inside the kernel
            pinfo[0].barrier()
            peer_ptr = pinfo[1]
        d_in, d_weight, d_bias = ops.layernorm_bwd(inputs, self.weight.data, mean, var, d_out, self.eps,
                                                   want_bf16=rt.is_bf16(), bf16_only=gemm_only and rt.is_bf16(),
                                                   want_colsum=gemm_only, dy_peer=peer_ptr)
        self.update_parameter("weight", gradient=d_weight)
        self.update_parameter("bias", gradient=d_bias)
        return d_in
