"""Layer-API parity on the GPU: the same constructor / forward / backward calls the reference's
own tests make (tests/test_linear.py, test_mlp.py, test_attention.py, test_normalization.py,
test_embedding.py, test_loss.py, test_transformer_block.py), checked against golden vectors
recorded from the unmodified reference.  fp32-accurate mode: rel 1e-4 norm-wise."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import numpitron_b200 as nb
    from numpitron_b200 import distributed as npdist
    from numpitron_b200.nn import (MLP, Attention, InputEmbedding, LayerNorm, Linear, OutputEmbedding,
                                   PositionalEncoding, TransformerBlock, softmax_cross_entropy)
    from numpitron_b200.runtime import to_numpy


@pytest.fixture(autouse=True)
def _world():
    npdist.init(tp_size=npdist.world_size())
    nb.set_precision("fp32")
    yield
    nb.set_precision("fp32")


def close(got, want, rtol=1e-4):
    got, want = to_numpy(got), np.asarray(want)
    np.testing.assert_allclose(got, want, rtol=rtol, atol=rtol * np.abs(want).max())


def test_linear(golden_layers):
    G = golden_layers
    lin = Linear(16, 24, weight_init="zeros")
    lin.update_parameter("weight", data=G["linear/w"])
    lin.update_parameter("bias", data=G["linear/b"])
    out = lin(G["linear/x"])
    assert isinstance(out, np.ndarray) and out.shape == (2, 8, 24)
    close(out, G["linear/y"], 1e-5)
    close(lin.backward(G["linear/dy"]), G["linear/dx"], 1e-5)
    close(lin.weight.gradient, G["linear/dw"], 1e-5)
    close(lin.bias.gradient, G["linear/db"], 1e-5)
    assert "inputs" not in lin.ctx  # popped by backward, like the reference


def test_mlp(golden_layers):
    G = golden_layers
    mlp = MLP(32, 128, 32, weight_init="zeros")
    mlp.column_linear.update_parameter("weight", data=G["mlp/w1"])
    mlp.column_linear.update_parameter("bias", data=G["mlp/b1"])
    mlp.row_linear.update_parameter("weight", data=G["mlp/w2"])
    mlp.row_linear.update_parameter("bias", data=G["mlp/b2"])
    close(mlp(G["mlp/x"]), G["mlp/y"], 1e-5)
    close(mlp.backward(G["mlp/dy"]), G["mlp/dx"], 1e-5)
    close(mlp.column_linear.weight.gradient, G["mlp/dw1"], 1e-5)
    close(mlp.column_linear.bias.gradient, G["mlp/db1"], 1e-5)
    close(mlp.row_linear.weight.gradient, G["mlp/dw2"], 1e-5)
    close(mlp.row_linear.bias.gradient, G["mlp/db2"], 1e-5)


def test_attention(golden_layers):
    G = golden_layers
    att = Attention(32, 4, 8, weight_init="zeros")
    att.qkv_projection.update_parameter("weight", data=G["att/wqkv"])
    att.out_projection.update_parameter("weight", data=G["att/wout"])
    close(att(G["att/x"]), G["att/y"], 1e-5)
    close(att.ctx["attention_weights"], G["att/p"], 1e-5)
    close(att.backward(G["att/dy"]), G["att/dx"], 2e-5)
    close(att.qkv_projection.weight.gradient, G["att/dwqkv"], 2e-5)   # element-wise, not just .sum()
    close(att.out_projection.weight.gradient, G["att/dwout"], 2e-5)


def test_layernorm(golden_layers):
    G = golden_layers
    ln = LayerNorm(32, eps=1e-5, weight_init="zeros")
    ln.update_parameter("weight", data=G["ln/w"])
    ln.update_parameter("bias", data=G["ln/b"])
    close(ln(G["ln/x"]), G["ln/y"], 1e-5)
    close(ln.backward(G["ln/dy"]), G["ln/dx"], 2e-5)
    close(ln.weight.gradient, G["ln/dw"], 1e-5)
    close(ln.bias.gradient, G["ln/db"], 1e-5)


def test_embeddings_and_quirk_q7(golden_layers):
    G = golden_layers
    emb = InputEmbedding(32, 17, weight_init="zeros")
    emb.update_parameter("embedding", data=G["emb/E"])
    out = OutputEmbedding(emb)
    emb.scatter()
    out.scatter()
    assert npdist.get_var("embedding_chunk_end") == 16
    y = emb(G["emb/tokens"])
    np.testing.assert_array_equal(y, G["emb/y"])               # bit-exact index work
    pe = PositionalEncoding(32, 16)
    np.testing.assert_array_equal(to_numpy(pe.encoding), G["emb/pos"])
    np.testing.assert_array_equal(pe(y), G["emb/y_pos"])
    close(out(G["emb/out_x"]), G["emb/logits"], 1e-5)
    close(out.backward(G["emb/dlogits"]), G["emb/out_dx"], 1e-5)
    close(emb.embedding.gradient, G["emb/dE_gemm"], 1e-5)
    # bit-exact ordered scatter-add given the same starting gradient
    emb.update_parameter("embedding", gradient=G["emb/dE_gemm"].copy())
    emb(G["emb/tokens"])
    assert emb.backward(G["emb/dy"]) is None
    np.testing.assert_array_equal(to_numpy(emb.embedding.gradient), G["emb/dE"])


def test_loss(golden_layers):
    G = golden_layers
    npdist.set_var("embedding_chunk_start", 0)
    npdist.set_var("embedding_chunk_end", 16)
    loss, d = softmax_cross_entropy(G["loss/logits"], G["loss/labels"])
    close(loss, G["loss/loss"], 1e-6)
    close(d, G["loss/dlogits"], 1e-6)


def _golden_block(G):
    blk = TransformerBlock(32, 4, weight_init="zeros")
    blk.norm1.update_parameter("weight", data=G["block/n1w"])
    blk.norm2.update_parameter("weight", data=G["block/n2w"])
    blk.attention.qkv_projection.update_parameter("weight", data=G["block/wqkv"])
    blk.attention.out_projection.update_parameter("weight", data=G["block/wout"])
    blk.mlp.column_linear.update_parameter("weight", data=G["block/w1"])
    blk.mlp.row_linear.update_parameter("weight", data=G["block/w2"])
    return blk


def test_transformer_block_fp32(golden_layers):
    G = golden_layers
    blk = _golden_block(G)
    close(blk(G["block/x"]), G["block/y"], 1e-5)
    close(blk.backward(G["block/dy"]), G["block/dx"], 1e-4)
    close(blk.attention.qkv_projection.weight.gradient, G["block/dwqkv"], 1e-4)
    close(blk.mlp.column_linear.weight.gradient, G["block/dw1"], 1e-4)
    close(blk.norm1.weight.gradient, G["block/dn1w"], 1e-4)
    close(blk.norm2.weight.gradient, G["block/dn2w"], 1e-4)


def test_transformer_block_bf16(golden_layers):
    """bf16 operands on tcgen05, fp32 accumulate / norms: looser, operand-rounding-limited."""
    G = golden_layers
    nb.set_precision("bf16")
    blk = _golden_block(G)
    close(blk(G["block/x"]), G["block/y"], 2e-2)
    close(blk.backward(G["block/dy"]), G["block/dx"], 5e-2)
    close(blk.mlp.column_linear.weight.gradient, G["block/dw1"], 5e-2)
