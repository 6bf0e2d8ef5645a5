"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Needs `/root/reference` (only present in the build container) and the `mpi4py` stand-in in
`oracle/mpi4py_stub`.  The outputs (`*.npz`) are committed; this script is committed so they
can be regenerated.  Usage (from the repo root):

    python tests/golden/make_golden.py layers        # single rank, per-layer + tiny train steps
    python tests/golden/make_golden.py world 2 1     # tp=2 dp=1 tiny train steps (spawns ranks)
    python tests/golden/make_golden.py world 2 2     # tp=2 dp=2
    python tests/golden/make_golden.py curve         # config-1 shape, 100 steps, B=2 (slow)
    python tests/golden/make_golden.py checkpoint    # a checkpoint file written by the reference + resumed step
    python tests/golden/make_golden.py loader        # batches of the reference DataLoader on a synthetic token file

Nothing in `tests/` reads `/root/reference` at test time.
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT / "oracle" / "mpi4py_stub"))
sys.path.insert(0, "/root/reference/src")

import numpy as np  # noqa: E402

TINY = dict(d_model=32, n_heads=4, n_layers=2, vocab_size=17, seq_len=16)
TINY_BATCH = 4


def tiny_batches(n_steps, vocab, batch, seq):
    rng = np.random.default_rng(7)
    data = rng.integers(0, vocab, 4096).astype(np.uint16)
    out = []
    for _ in range(n_steps):
        start = rng.integers(0, len(data) - 1 - seq, size=batch)[:, None]
        out.append((data[start + np.arange(seq)], data[start + 1 + np.arange(seq)]))
    return out


def layers():
    from numpitron import distributed as npdist
    from numpitron.nn import (MLP, Attention, InputEmbedding, LayerNorm, Linear, OutputEmbedding,
                              PositionalEncoding, Transformer, TransformerBlock, softmax_cross_entropy)
    from numpitron.nn.activation import ReLU, Softmax
    from numpitron.optimizer import Adam

    npdist.init(tp_size=1, dp_size=1)
    G = {}
    f32 = np.float32

    # ---- Linear --------------------------------------------------------------------------
    rng = np.random.default_rng(100)
    lin = Linear(16, 24, rng=rng)
    lin.update_parameter("bias", data=rng.standard_normal(24).astype(f32))
    x = rng.standard_normal((2, 8, 16)).astype(f32)
    dy = rng.standard_normal((2, 8, 24)).astype(f32)
    G["linear/w"], G["linear/b"], G["linear/x"], G["linear/dy"] = lin.weight.data, lin.bias.data, x, dy
    G["linear/y"] = lin(x)
    G["linear/dx"] = lin.backward(dy)
    G["linear/dw"], G["linear/db"] = lin.weight.gradient, lin.bias.gradient

    # ---- ReLU / Softmax ------------------------------------------------------------------
    relu = ReLU()
    x = rng.standard_normal((2, 8, 24)).astype(f32)
    dy = rng.standard_normal((2, 8, 24)).astype(f32)
    G["relu/x"], G["relu/dy"], G["relu/y"] = x, dy, relu(x)
    G["relu/dx"] = relu.backward(dy)
    sm = Softmax()
    s = 12
    x = rng.standard_normal((2, 3, s, s)).astype(f32)
    x = np.where(np.tri(s, s, dtype=bool)[None, None], x, -np.inf).astype(f32)
    dy = rng.standard_normal((2, 3, s, s)).astype(f32)
    G["softmax/x"], G["softmax/dy"], G["softmax/y"] = x, dy, sm(x)
    G["softmax/dx"] = sm.backward(dy)

    # ---- LayerNorm (inputs with a large mean, like the model's: SURVEY Q12) --------------------
    ln = LayerNorm(32, rng=rng)
    ln.update_parameter("bias", data=(rng.standard_normal(32) * 0.1).astype(f32))
    x = (1.5 + 0.03 * rng.standard_normal((2, 8, 32))).astype(f32)
    dy = rng.standard_normal((2, 8, 32)).astype(f32)
    G["ln/w"], G["ln/b"], G["ln/x"], G["ln/dy"] = ln.weight.data, ln.bias.data, x, dy
    G["ln/y"] = ln(x)
    G["ln/dx"] = ln.backward(dy)
    G["ln/dw"], G["ln/db"] = ln.weight.gradient, ln.bias.gradient

    # ---- MLP --------------------------------------------------------------------------------
    mlp = MLP(32, 128, 32, rng=rng)
    mlp.column_linear.update_parameter("bias", data=(rng.standard_normal(128) * 0.05 - 0.3).astype(f32))
    mlp.row_linear.update_parameter("bias", data=(rng.standard_normal(32) * 0.05).astype(f32))
    x = rng.standard_normal((2, 8, 32)).astype(f32)
    dy = rng.standard_normal((2, 8, 32)).astype(f32)
    G["mlp/w1"], G["mlp/b1"] = mlp.column_linear.weight.data, mlp.column_linear.bias.data
    G["mlp/w2"], G["mlp/b2"] = mlp.row_linear.weight.data, mlp.row_linear.bias.data
    G["mlp/x"], G["mlp/dy"], G["mlp/y"] = x, dy, mlp(x)
    G["mlp/dx"] = mlp.backward(dy)
    G["mlp/dw1"], G["mlp/db1"] = mlp.column_linear.weight.gradient, mlp.column_linear.bias.gradient
    G["mlp/dw2"], G["mlp/db2"] = mlp.row_linear.weight.gradient, mlp.row_linear.bias.gradient

    # ---- Attention ----------------------------------------------------------------------------
    att = Attention(32, 4, 8, rng=rng, scale=0.5)
    x = rng.standard_normal((2, 16, 32)).astype(f32)
    dy = rng.standard_normal((2, 16, 32)).astype(f32)
    G["att/wqkv"], G["att/wout"] = att.qkv_projection.weight.data, att.out_projection.weight.data
    G["att/x"], G["att/dy"], G["att/y"] = x, dy, att(x)
    G["att/p"] = att.ctx["attention_weights"]
    G["att/dx"] = att.backward(dy)
    G["att/dwqkv"], G["att/dwout"] = att.qkv_projection.weight.gradient, att.out_projection.weight.gradient

    # ---- Embeddings (tp=1 but scattered, so the vocab-1 quirk Q7 is active) -------------------
    emb = InputEmbedding(32, 17, rng=rng)
    out = OutputEmbedding(emb)
    emb.scatter()
    out.scatter()
    tok = rng.integers(0, 17, (3, 16)).astype(np.uint16)
    tok[0, :3] = 16
    pe = PositionalEncoding(32, 16)
    G["emb/E"], G["emb/tokens"], G["emb/pos"] = emb.embedding.data, tok, pe.encoding
    e = emb(tok.copy())
    G["emb/y"] = e
    G["emb/y_pos"] = pe(e)
    xo = rng.standard_normal((3, 16, 32)).astype(f32)
    G["emb/out_x"] = xo
    G["emb/logits"] = out(xo)
    dl = rng.standard_normal((3, 16, 17)).astype(f32)
    G["emb/dlogits"] = dl
    G["emb/out_dx"] = out.backward(dl)
    G["emb/dE_gemm"] = emb.embedding.gradient.copy()
    dlast = rng.standard_normal((3, 16, 32)).astype(f32)
    G["emb/dy"] = dlast
    emb.backward(dlast)
    G["emb/dE"] = emb.embedding.gradient

    # ---- loss (chunk bounds still (0,16) from the embedding above) ------------------------------
    logits = (rng.standard_normal((3, 16, 17)) * 3).astype(f32)
    labels = rng.integers(0, 17, (3, 16)).astype(np.uint16)
    labels[1, :2] = 16
    G["loss/logits"], G["loss/labels"] = logits, labels
    G["loss/loss"], G["loss/dlogits"] = softmax_cross_entropy(logits, labels.copy())

    # ---- TransformerBlock ---------------------------------------------------------------------------
    blk = TransformerBlock(32, 4, rng=rng)
    x = (1.0 + rng.standard_normal((2, 16, 32))).astype(f32)
    dy = rng.standard_normal((2, 16, 32)).astype(f32)
    G["block/x"], G["block/dy"] = x, dy
    G["block/n1w"], G["block/n2w"] = blk.norm1.weight.data, blk.norm2.weight.data
    G["block/wqkv"], G["block/wout"] = blk.attention.qkv_projection.weight.data, blk.attention.out_projection.weight.data
    G["block/w1"], G["block/w2"] = blk.mlp.column_linear.weight.data, blk.mlp.row_linear.weight.data
    G["block/y"] = blk(x)
    G["block/dx"] = blk.backward(dy)
    G["block/dwqkv"] = blk.attention.qkv_projection.weight.gradient
    G["block/dw1"] = blk.mlp.column_linear.weight.gradient
    G["block/dn1w"], G["block/dn2w"] = blk.norm1.weight.gradient, blk.norm2.weight.gradient

    # ---- Adam: one parameter, three updates ------------------------------------------------------------
    lin = Linear(8, 8, rng=rng)
    class _M:  # minimal Model-like holder for Adam's tree walk
        pass
    from numpitron.nn.model import Model
    m = Model()
    m.add_layer("lin", lin)
    opt = Adam(m, 1e-3, 0.9, 0.99)
    G["adam/p0"] = lin.weight.data.copy()
    for i in range(3):
        g = (rng.standard_normal((8, 8)) * 10.0 ** (-3 * i)).astype(f32)
        G[f"adam/g{i}"] = g
        lin.update_parameter("weight", gradient=g)
        lin.update_parameter("bias", gradient=np.zeros(8, f32))
        opt.step()
        G[f"adam/p{i + 1}"] = lin.weight.data.copy()
    G["adam/m"], G["adam/v"] = opt.parameters["lin"]["weight"].momentum.data, opt.parameters["lin"]["weight"].velocity.data

    # ---- tiny Transformer: 3 train steps (train_shakespeare.py:14-21) -----------------------------------
    rng = np.random.default_rng(42)
    t = Transformer(**TINY, rng=rng)
    opt = Adam(t, 1e-4, 0.9, 0.99)
    t.scatter()
    opt.scatter()
    batches = tiny_batches(3, TINY["vocab_size"], TINY_BATCH, TINY["seq_len"])
    for i, (x, y) in enumerate(batches):
        G[f"tiny/x{i}"], G[f"tiny/y{i}"] = x, y
        logits = t(x.copy())
        loss, d = softmax_cross_entropy(logits, y.copy())
        t.backward(d)
        if i == 0:
            G["tiny/logits0"], G["tiny/dlogits0"] = logits, d
            G["tiny/dE0"] = t.input_embedding.embedding.gradient.copy()
            G["tiny/dwqkv0_block1"] = t.block_1.attention.qkv_projection.weight.gradient.copy()
            G["tiny/dw1_0_block1"] = t.block_1.mlp.column_linear.weight.gradient.copy()
            G["tiny/dn2w0_block1"] = t.block_1.norm2.weight.gradient.copy()
        opt.step()
        G[f"tiny/loss{i}"] = loss
    G["tiny/E_final"] = t.input_embedding.embedding.data
    G["tiny/wqkv_final_block1"] = t.block_1.attention.qkv_projection.weight.data
    np.savez_compressed(HERE / "layers_tp1.npz", **G)
    print("wrote layers_tp1.npz", sum(v.nbytes for v in G.values()) / 1e6, "MB raw")


def world_rank_main(tp, dp):
    """Runs inside each spawned rank (WORLD_SIZE/RANK in env)."""
    from numpitron import distributed as npdist
    from numpitron.nn import Transformer, softmax_cross_entropy
    from numpitron.optimizer import Adam

    npdist.init(tp_size=tp, dp_size=dp)
    rng = np.random.default_rng(42)
    t = Transformer(**TINY, rng=rng)
    opt = Adam(t, 1e-4, 0.9, 0.99)
    t.scatter()
    opt.scatter()
    G = {}
    batch = TINY_BATCH
    for i, (x, y) in enumerate(tiny_batches(3, TINY["vocab_size"], batch, TINY["seq_len"])):
        if dp > 1:  # loader.py:56-80 — Scatterv along batch within the DP group
            x = np.array_split(x, dp, axis=0)[npdist.data_parallel_rank()]
            y = np.array_split(y, dp, axis=0)[npdist.data_parallel_rank()]
        logits = t(x.copy())
        loss, d = softmax_cross_entropy(logits, y.copy())
        t.backward(d)
        if i == 0:
            G["logits0"] = logits
            G["dE0"] = t.input_embedding.embedding.gradient.copy()
        opt.step()
        G[f"loss{i}"] = loss
    G["E_final"] = t.input_embedding.embedding.data
    G["w1_final_block0"] = t.block_0.mlp.column_linear.weight.data
    np.savez_compressed(f"/tmp/golden_world_{tp}_{dp}_rank{npdist.world_rank()}.npz", **G)


def world(tp, dp):
    n = tp * dp
    procs = []
    for r in range(n):
        env = dict(os.environ, WORLD_SIZE=str(n), RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT="29533", OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, __file__, "_rank", str(tp), str(dp)], env=env))
    for p in procs:
        assert p.wait(timeout=600) == 0
    G = {}
    for r in range(n):
        with np.load(f"/tmp/golden_world_{tp}_{dp}_rank{r}.npz") as z:
            for k in z.files:
                G[f"rank{r}/{k}"] = z[k]
    np.savez_compressed(HERE / f"tiny_tp{tp}_dp{dp}.npz", **G)
    print(f"wrote tiny_tp{tp}_dp{dp}.npz")


def curve():
    """Config-1 shape (L6 d384 H6 S256 V65), B=2, 100 steps, the literal reference."""
    from numpitron import distributed as npdist
    from numpitron.nn import Transformer, softmax_cross_entropy
    from numpitron.optimizer import Adam

    npdist.init(tp_size=1, dp_size=1)
    rng = np.random.default_rng(42)
    t = Transformer(d_model=384, n_heads=6, n_layers=6, vocab_size=65, seq_len=256, rng=rng)
    opt = Adam(t, 1e-4, 0.9, 0.99)
    t.scatter()
    opt.scatter()
    data = np.random.default_rng(1234).integers(0, 65, 2_000_000).astype(np.uint16)
    brng = np.random.default_rng(99)
    losses = []
    n_steps = int(os.environ.get("CURVE_STEPS", "100"))
    for i in range(n_steps):
        start = brng.integers(0, len(data) - 1 - 256, size=2)[:, None]
        x, y = data[start + np.arange(256)], data[start + 1 + np.arange(256)]
        logits = t(x.copy())
        loss, d = softmax_cross_entropy(logits, y.copy())
        t.backward(d)
        opt.step()
        losses.append(loss.mean())
        print(i, losses[-1], flush=True)
        np.savez_compressed(HERE / "curve_c1_b2.npz", losses=np.array(losses, dtype=np.float32),
                            E_final=t.input_embedding.embedding.data)


LOADER_TOKENS = dict(seed=1234, vocab=17, n=5000)   # the synthetic uint16 token file of the loader fixture
LOADER_CFG = dict(seq_len=16, batch_size=4, rng_seed=5)


def loader():
    """the first 6 batches and the epoch length of the reference's DataLoader
    (data/loader.py) on a synthetic token file that the test regenerates from LOADER_TOKENS."""
    import tempfile

    from numpitron import distributed as npdist
    from numpitron.data import DataLoader

    npdist.init(tp_size=1, dp_size=1)
    data = np.random.default_rng(LOADER_TOKENS["seed"]).integers(0, LOADER_TOKENS["vocab"], LOADER_TOKENS["n"]).astype(np.uint16)
    with tempfile.TemporaryDirectory() as tmp:
        path = Path(tmp) / "tokens.bin"
        data.tofile(path)
        dl = DataLoader(path, LOADER_CFG["seq_len"], LOADER_CFG["batch_size"], np.random.default_rng(LOADER_CFG["rng_seed"]))
        G = {"batches_per_epoch": np.int64(dl.batches_per_epoch)}
        n_epoch = 0
        for i, (x, y) in enumerate(dl.iter_epoch()):
            n_epoch += 1
            if i < 6:
                G[f"x{i}"], G[f"y{i}"] = np.array(x), np.array(y)
        G["batches_yielded_per_epoch"] = np.int64(n_epoch)
    np.savez_compressed(HERE / "loader_tp1.npz", **G)
    print("wrote loader_tp1.npz", int(G["batches_per_epoch"]), n_epoch)


def checkpoint():
    """the reference's own `checkpoint()` (train_shakespeare.py:30-41) writes the tiny model
    after 2 training steps; the file is reloaded the way `train()` does (:87-93) and step 3 is run from
    it.  Outputs: tiny_checkpoint.npy (the pickled state, as the reference wrote it), tiny_checkpoint_resaved.npy
    (the same state loaded and saved again by the reference) and
    tiny_checkpoint_resume.npz (batch, loss and embedding of the resumed step)."""
    sys.path.insert(0, "/root/reference")
    import train_shakespeare as ref_script
    from numpitron import distributed as npdist
    from numpitron.nn import Transformer, softmax_cross_entropy
    from numpitron.optimizer import Adam

    npdist.init(tp_size=1, dp_size=1)
    t = Transformer(**TINY, rng=np.random.default_rng(42))
    opt = Adam(t, 1e-4, 0.9, 0.99)
    t.scatter()
    opt.scatter()
    batches = tiny_batches(3, TINY["vocab_size"], TINY_BATCH, TINY["seq_len"])
    loss = None
    for x, y in batches[:2]:
        logits = t(x.copy())
        loss, d = softmax_cross_entropy(logits, y.copy())
        t.backward(d)
        opt.step()
    path = HERE / "tiny_checkpoint.npy"
    ref_script.checkpoint(t, opt, 2, float(loss.mean()), path)
    # resume exactly like train_shakespeare.py:87-93
    state = np.load(path, allow_pickle=True)[()]
    t2 = Transformer.from_dict(state["model_parameters"])
    opt2 = Adam.from_dict(state["optimizer_parameters"], t2)
    t2.scatter()
    opt2.scatter()
    # what the reference itself writes when it saves the reloaded state again (from_dict rebuilds the layers
    # with weight_init="zeros", so those settings differ from the first file)
    ref_script.checkpoint(t2, opt2, state["epoch"], state["validation_loss"], HERE / "tiny_checkpoint_resaved.npy")
    x, y = batches[2]
    logits = t2(x.copy())
    loss3, d = softmax_cross_entropy(logits, y.copy())
    t2.backward(d)
    opt2.step()
    np.savez_compressed(HERE / "tiny_checkpoint_resume.npz", x=x, y=y, loss=loss3, epoch=state["epoch"],
                        validation_loss=state["validation_loss"], E_after=t2.input_embedding.embedding.data,
                        wqkv_after_block1=t2.block_1.attention.qkv_projection.weight.data)
    print("wrote tiny_checkpoint.npy", path.stat().st_size, "bytes; resumed loss", float(loss3.mean()))


if __name__ == "__main__":
    what = sys.argv[1]
    if what == "layers":
        layers()
    elif what == "world":
        world(int(sys.argv[2]), int(sys.argv[3]))
    elif what == "_rank":
        world_rank_main(int(sys.argv[2]), int(sys.argv[3]))
    elif what == "curve":
        curve()
    elif what == "checkpoint":
        checkpoint()
    elif what == "loader":
        loader()
