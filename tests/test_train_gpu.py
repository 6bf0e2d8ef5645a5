"""End-to-end train-step parity on the GPU (train_shakespeare.py:14-21 of the reference).

* tiny model, 3 steps: losses / logits / gradients / updated weights against golden vectors of
  the unmodified reference;
* config-1 shape (L6 d384 H6 S256 V65), B=2: the loss curve against the literal reference's
  first steps (tests/golden/curve_c1_b2.npz) — rel 1e-3 (the reference's own
  fp32-vs-fp64 noise is 6e-4 relative late in the curve);
* bf16 mode: loss within 2e-2 (north_star);
* the CUDA-graph replay of the step gives the same numbers as the eager calls;
* size-independent properties at the full bench size.
"""
import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import numpitron_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import numpitron_b200 as nb
    from numpitron_b200 import distributed as npdist
    from numpitron_b200.nn import Transformer, softmax_cross_entropy
    from numpitron_b200.optimizer import Adam
    from numpitron_b200.runtime import to_numpy
    from numpitron_b200.train import GraphedTrainStep, train_step

TINY = dict(d_model=32, n_heads=4, n_layers=2, vocab_size=17, seq_len=16)
C1 = dict(d_model=384, n_heads=6, n_layers=6, vocab_size=65, seq_len=256)


@pytest.fixture(autouse=True)
def _world():
    npdist.init(tp_size=1, dp_size=1)
    nb.set_precision("fp32")
    yield
    nb.set_precision("fp32")


def build(cfg, seed=42, lr=1e-4):
    t = Transformer(**cfg, rng=np.random.default_rng(seed))
    opt = Adam(t, lr, 0.9, 0.99)
    t.scatter()
    opt.scatter()
    return t, opt


def close(got, want, rtol):
    got, want = to_numpy(got), np.asarray(want)
    np.testing.assert_allclose(got, want, rtol=rtol, atol=rtol * np.abs(want).max())


def test_weights_identical_to_reference_init(golden_layers):
    """Same rng call order as the reference -> identical random-init weights (bit-exact)."""
    t, _ = build(TINY)
    full = O.build_full_parameters(32, 4, 2, 17, np.random.default_rng(42))
    np.testing.assert_array_equal(to_numpy(t.input_embedding.embedding.data), full["embedding"])
    np.testing.assert_array_equal(to_numpy(t.block_1.mlp.row_linear.weight.data), full["block_1.row.weight"])
    np.testing.assert_array_equal(to_numpy(t.block_0.norm2.weight.data), full["block_0.norm2.weight"])


def test_tiny_three_steps_vs_golden(golden_layers):
    G = golden_layers
    t, opt = build(TINY)
    for i in range(3):
        x, y = G[f"tiny/x{i}"], G[f"tiny/y{i}"]
        logits = t(x)
        loss, d = softmax_cross_entropy(logits, y)
        t.backward(d)
        if i == 0:
            close(logits, G["tiny/logits0"], 1e-5)
            close(d, G["tiny/dlogits0"], 1e-4)
            close(t.input_embedding.embedding.gradient, G["tiny/dE0"], 1e-4)
            close(t.block_1.attention.qkv_projection.weight.gradient, G["tiny/dwqkv0_block1"], 1e-4)
            close(t.block_1.mlp.column_linear.weight.gradient, G["tiny/dw1_0_block1"], 1e-4)
            close(t.block_1.norm2.weight.gradient, G["tiny/dn2w0_block1"], 1e-4)
        opt.step()
        close(loss, G[f"tiny/loss{i}"], 1e-5)
        assert t.block_0.norm1.weight.gradient is None  # consumed by Adam, like the reference
    close(t.input_embedding.embedding.data, G["tiny/E_final"], 1e-5)
    close(t.block_1.attention.qkv_projection.weight.data, G["tiny/wqkv_final_block1"], 1e-5)


def _curve_batches(n):
    data = O.synthetic_tokens(65)
    brng = np.random.default_rng(99)
    for _ in range(n):
        yield O.draw_batch(data, brng, 2, 256)


def _golden_curve():
    path = GOLDEN / "curve_c1_b2.npz"
    if not path.exists():
        pytest.skip("curve fixture not generated")
    with np.load(path) as z:
        return z["losses"]


@pytest.mark.parametrize("precision,rtol", [("fp32", 1e-3), ("bf16", 2e-2)])
def test_config1_loss_curve_vs_literal_reference(precision, rtol):
    want = _golden_curve()
    n = len(want)
    nb.set_precision(precision)
    t, opt = build(C1)
    got = [train_step(t, opt, x, y) for x, y in _curve_batches(n)]
    err = np.abs(np.array(got) - want) / np.abs(want)
    print(f"{precision}: {n} steps, loss {got[0]:.4f}->{got[-1]:.4f} (ref {want[0]:.4f}->{want[-1]:.4f}), "
          f"max rel err {err.max():.2e}")
    assert err.max() < rtol, f"loss curve deviates: {err.max()}"


def test_config1_vs_fast_oracle_longer_run():
    """40 steps against the oracle's fast mode (compact softmax backward, BLAS matmul)."""
    n = 40
    t, opt = build(C1)
    orc = O.OracleTransformer(384, 6, 6, 65, 256, rng=np.random.default_rng(42), literal=False)
    worst = 0.0
    for x, y in _curve_batches(n):
        a = train_step(t, opt, x, y)
        b = float(orc.train_step(x, y)[0])
        worst = max(worst, abs(a - b) / abs(b))
    print(f"{n} steps vs fast oracle: final loss {a:.4f} / {b:.4f}, max rel err {worst:.2e}")
    assert worst < 1e-3
    close(t.input_embedding.embedding.data, orc.ranks[0][0].params["embedding"], 1e-3)


def test_graph_replay_matches_eager():
    rng = np.random.default_rng(3)
    xs = [rng.integers(0, 17, (4, 16)).astype(np.uint16) for _ in range(6)]
    ys = [rng.integers(0, 17, (4, 16)).astype(np.uint16) for _ in range(6)]
    t1, o1 = build(TINY)
    eager = [train_step(t1, o1, x, y) for x, y in zip(xs, ys)]
    t2, o2 = build(TINY)
    g = GraphedTrainStep(t2, o2, 4, 16, warmup=2)
    graphed = [g.step(x, y) for x, y in zip(xs, ys)]
    assert g.graph is not None
    np.testing.assert_array_equal(np.float32(eager), np.float32(graphed))
    np.testing.assert_array_equal(to_numpy(t1.input_embedding.embedding.data),
                                  to_numpy(t2.input_embedding.embedding.data))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_full_size_properties(precision):
    """Bench-size batch (64 x 256 tokens): loss finite and decreasing, every parameter got a
    gradient and was consumed, token V-1 contributes a zero embedding row (Q7), the step is
    deterministic (two identical runs agree bit for bit)."""
    nb.set_precision(precision)
    data = O.synthetic_tokens(65)

    def run():
        t, opt = build(C1)
        brng = np.random.default_rng(5)
        losses = []
        for _ in range(4):
            x, y = O.draw_batch(data, brng, 64, 256)
            losses.append(train_step(t, opt, x, y))
        return losses, to_numpy(t.input_embedding.embedding.data)

    l1, e1 = run()
    l2, e2 = run()
    assert np.all(np.isfinite(l1)) and l1[-1] < l1[0]
    assert 20 < l1[0] < 40  # initial loss ~31 at this shape with the reference's init
    np.testing.assert_array_equal(np.float32(l1), np.float32(l2))
    np.testing.assert_array_equal(e1, e2)


def _tree_equal(a, b, path="state"):
    """Two checkpoint states are the same tree: same keys in the same order, same leaf types, dtypes,
    shapes and (bit-exact) values."""
    if isinstance(a, dict):
        assert isinstance(b, dict) and list(a.keys()) == list(b.keys()), (path, list(a), list(b) if isinstance(b, dict) else b)
        for k in a:
            _tree_equal(a[k], b[k], f"{path}[{k!r}]")
    elif isinstance(a, np.ndarray):
        assert isinstance(b, np.ndarray) and a.dtype == b.dtype and a.shape == b.shape, (path, a.dtype, getattr(b, "dtype", b))
        np.testing.assert_array_equal(a, b, err_msg=path)
    else:
        assert type(a) is type(b) and a == b, (path, a, b)


def test_checkpoint_interop_with_reference(tmp_path):
    """SURVEY §8(f)1.  tests/golden/tiny_checkpoint.npy was written by the reference's own checkpoint()
    after 2 training steps (make_golden.py checkpoint).  (1) It loads here, and saving it again from
    here gives the tree the reference itself writes when it loads and re-saves that file
    (tiny_checkpoint_resaved.npy) — every key, dtype, shape and value — so the reference can read our
    files.  (2) The step resumed from it reproduces the reference's own resumed step."""
    from numpitron_b200.checkpoint import load_checkpoint, save_checkpoint

    src = GOLDEN / "tiny_checkpoint.npy"
    want = np.load(src, allow_pickle=True)[()]
    t, opt, epoch, vloss = load_checkpoint(src)
    assert epoch == want["epoch"] and vloss == want["validation_loss"]
    assert opt.timestep == want["optimizer_parameters"]["timestep"]
    t.scatter()
    opt.scatter()
    out = tmp_path / "roundtrip.npy"
    save_checkpoint(t, opt, epoch, vloss, out)
    _tree_equal(np.load(GOLDEN / "tiny_checkpoint_resaved.npy", allow_pickle=True)[()], np.load(out, allow_pickle=True)[()])

    with np.load(GOLDEN / "tiny_checkpoint_resume.npz") as z:
        R = {k: z[k] for k in z.files}
    logits = t(R["x"])
    loss, d = softmax_cross_entropy(logits, R["y"])
    t.backward(d)
    opt.step()
    close(loss, R["loss"], 1e-5)
    close(t.input_embedding.embedding.data, R["E_after"], 1e-5)
    close(t.block_1.attention.qkv_projection.weight.data, R["wqkv_after_block1"], 1e-5)


def test_device_dataloader_matches_reference(tmp_path):
    """SURVEY §8(f)2.  The device-resident DataLoader (token file in HBM, start indices from the same NumPy
    generator, windows cut by npt_window_gather) yields the reference DataLoader's batches bit for bit
    (tests/golden/loader_tp1.npz, written by the reference's own loader), the same number of batches per
    epoch, and its batches feed train_step directly."""
    from numpitron_b200.data import DataLoader

    with np.load(GOLDEN / "loader_tp1.npz") as z:
        G = {k: z[k] for k in z.files}
    data = np.random.default_rng(1234).integers(0, 17, 5000).astype(np.uint16)   # make_golden.LOADER_TOKENS
    path = tmp_path / "tokens.bin"
    data.tofile(path)
    dl = DataLoader(path, 16, 4, np.random.default_rng(5))
    assert dl.batches_per_epoch == int(G["batches_per_epoch"])
    n = 0
    for i, (x, y) in enumerate(dl.iter_epoch()):
        n += 1
        assert x.is_cuda and x.dtype == torch.uint16 and tuple(x.shape) == (4, 16)
        if i < 6:
            np.testing.assert_array_equal(x.view(torch.int16).cpu().numpy().view(np.uint16), G[f"x{i}"])
            np.testing.assert_array_equal(y.view(torch.int16).cpu().numpy().view(np.uint16), G[f"y{i}"])
    assert n == int(G["batches_yielded_per_epoch"])
    # the device batches are what the layers take: one train step from them equals the step from host arrays
    t1, o1 = build(TINY)
    t2, o2 = build(TINY)
    dl2 = DataLoader(path, 16, 4, np.random.default_rng(5))
    xd, yd = dl2.next_batch()
    l_dev = train_step(t1, o1, xd, yd)
    l_host = train_step(t2, o2, G["x0"], G["y0"])
    assert l_dev == l_host


def test_sampling_matches_reference():
    """SURVEY §8(f)3.  Tokens generated from the reference-written checkpoint with the reference's samplers
    (tests/golden/sample_tp1.npz: 24 tokens from a 1-token prompt, so the 16-token context window slides)
    — greedy, and temperature sampling with the same NumPy generator — are reproduced token for token in
    the fp32 mode."""
    from numpitron_b200.checkpoint import load_checkpoint
    from numpitron_b200.sample import generate

    with np.load(GOLDEN / "sample_tp1.npz") as z:
        G = {k: z[k] for k in z.files}
    t, _, _, _ = load_checkpoint(GOLDEN / "tiny_checkpoint.npy")
    t.scatter()
    assert generate(t, [3], 24, None, None, 17) == G["greedy"].tolist()
    for name, temp in (("basic_t1", 1.0), ("basic_t40", 40.0)):
        assert generate(t, [3], 24, temp, np.random.default_rng(42), 17) == G[name].tolist(), name


def test_graph_survives_checkpoint(tmp_path):
    """A reference-style loop checkpoints between training steps (train_shakespeare.py:118-126):
    save_checkpoint() gathers and re-scatters the model and the optimizer, which replaces every parameter /
    moment buffer.  The captured graph must notice (nn.core.storage_generation) and re-capture on the new
    storage instead of training the orphaned buffers: graphed == eager, before and after the checkpoint."""
    from numpitron_b200.checkpoint import save_checkpoint

    rng = np.random.default_rng(11)
    xs = [rng.integers(0, 17, (4, 16)).astype(np.uint16) for _ in range(10)]
    ys = [rng.integers(0, 17, (4, 16)).astype(np.uint16) for _ in range(10)]
    t1, o1 = build(TINY)
    t2, o2 = build(TINY)
    g = GraphedTrainStep(t2, o2, 4, 16, warmup=2)
    eager, graphed = [], []
    for i, (x, y) in enumerate(zip(xs, ys)):
        if i == 5:
            save_checkpoint(t1, o1, 0, 0.0, tmp_path / "a.npy")
            save_checkpoint(t2, o2, 0, 0.0, tmp_path / "b.npy")
            assert g.graph is not None   # it was captured before the checkpoint
        eager.append(train_step(t1, o1, x, y))
        graphed.append(g.step(x, y))
    assert g.graph is not None   # ... and re-captured after it
    np.testing.assert_array_equal(np.float32(eager), np.float32(graphed))
    np.testing.assert_array_equal(to_numpy(t1.input_embedding.embedding.data), to_numpy(t2.input_embedding.embedding.data))
    np.testing.assert_array_equal(to_numpy(t1.block_1.mlp.row_linear.weight.data), to_numpy(t2.block_1.mlp.row_linear.weight.data))
    a = np.load(tmp_path / "a.npy", allow_pickle=True)[()]
    b = np.load(tmp_path / "b.npy", allow_pickle=True)[()]
    _tree_equal(a, b)


def test_loader_feeds_graphed_step_without_host_sync(tmp_path):
    """DataLoader.next_batch launches its gather on the caller's stream; load_resident copies on the step's
    own stream.  No host synchronisation in between: the stream ordering alone must make the batches arrive."""
    from numpitron_b200.data import DataLoader

    data = np.random.default_rng(1234).integers(0, 17, 50000).astype(np.uint16)
    path = tmp_path / "tokens.bin"
    data.tofile(path)
    t1, o1 = build(TINY)
    t2, o2 = build(TINY)
    dl1 = DataLoader(path, 16, 4, np.random.default_rng(5))
    dl2 = DataLoader(path, 16, 4, np.random.default_rng(5))
    g = GraphedTrainStep(t2, o2, 4, 16, warmup=2)
    want, got = [], []
    for _ in range(8):
        x, y = dl1.next_batch()
        want.append(train_step(t1, o1, x, y))
    for _ in range(8):
        g.load_resident(*dl2.next_batch())
        g.step_resident_async()
        got.append(g.loss_dev.clone())   # same stream order as the step? no: read after the final sync below
    g.stream.synchronize()
    np.testing.assert_array_equal(to_numpy(t1.input_embedding.embedding.data), to_numpy(t2.input_embedding.embedding.data))
    assert float(g.loss_dev.item()) == want[-1]
