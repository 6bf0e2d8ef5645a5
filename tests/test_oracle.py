"""Pin the CPU oracle (oracle/numpitron_oracle.py) to golden vectors produced by the
UNMODIFIED reference (tests/golden/make_golden.py).  literal=True must be bit-exact."""
import numpy as np
import pytest

from conftest import GOLDEN, ROOT, load_golden
from oracle import numpitron_oracle as O

TINY = dict(d_model=32, n_heads=4, n_layers=2, vocab=17, seq_len=16)


def eq(a, b):
    np.testing.assert_array_equal(np.asarray(a), np.asarray(b))


def test_linear(golden_layers):
    G = golden_layers
    eq(O.linear_forward(G["linear/x"], G["linear/w"], G["linear/b"]), G["linear/y"])
    dx, dw, db = O.linear_backward(G["linear/x"], G["linear/w"], G["linear/dy"], True)
    eq(dx, G["linear/dx"]); eq(dw, G["linear/dw"]); eq(db, G["linear/db"])
    # fast path: same numbers to fp32 rounding
    np.testing.assert_allclose(O.linear_forward(G["linear/x"], G["linear/w"], G["linear/b"], literal=False),
                               G["linear/y"], rtol=1e-5, atol=1e-6)


def test_relu_softmax(golden_layers):
    G = golden_layers
    eq(O.relu_forward(G["relu/x"]), G["relu/y"])
    eq(O.relu_backward(G["relu/x"], G["relu/dy"]), G["relu/dx"])
    eq(O.softmax(G["softmax/x"]), G["softmax/y"])
    eq(O.softmax_backward_jacobian(G["softmax/y"], G["softmax/dy"]), G["softmax/dx"])
    np.testing.assert_allclose(O.softmax_backward_compact(G["softmax/y"], G["softmax/dy"]),
                               G["softmax/dx"], rtol=1e-5, atol=1e-7)


def test_layernorm(golden_layers):
    G = golden_layers
    y, ctx = O.layernorm_forward(G["ln/x"], G["ln/w"], G["ln/b"])
    eq(y, G["ln/y"])
    dx, dw, db = O.layernorm_backward(ctx, G["ln/w"], G["ln/dy"])
    eq(dx, G["ln/dx"]); eq(dw, G["ln/dw"]); eq(db, G["ln/db"])


def test_mlp(golden_layers):
    G = golden_layers
    a = O.linear_forward(G["mlp/x"], G["mlp/w1"], G["mlp/b1"])
    r = O.relu_forward(a)
    eq(O.linear_forward(r, G["mlp/w2"], G["mlp/b2"]), G["mlp/y"])
    dr, dw2, db2 = O.linear_backward(r, G["mlp/w2"], G["mlp/dy"], True)
    da = O.relu_backward(a, dr)
    dx, dw1, db1 = O.linear_backward(G["mlp/x"], G["mlp/w1"], da, True)
    eq(dx, G["mlp/dx"]); eq(dw1, G["mlp/dw1"]); eq(db1, G["mlp/db1"]); eq(dw2, G["mlp/dw2"]); eq(db2, G["mlp/db2"])


@pytest.mark.parametrize("literal", [True, False])
def test_attention(golden_layers, literal):
    G = golden_layers
    y, ctx = O.attention_forward(G["att/x"], G["att/wqkv"], G["att/wout"], 4, literal)
    dx, dwqkv, dwout = O.attention_backward(ctx, G["att/wqkv"], G["att/wout"], G["att/dy"], literal)
    if literal:
        eq(y, G["att/y"]); eq(ctx["p"], G["att/p"]); eq(dx, G["att/dx"]); eq(dwqkv, G["att/dwqkv"]); eq(dwout, G["att/dwout"])
    else:
        for got, want in ((y, "att/y"), (dx, "att/dx"), (dwqkv, "att/dwqkv"), (dwout, "att/dwout")):
            w = G[want]
            np.testing.assert_allclose(got, w, rtol=1e-5, atol=1e-5 * np.abs(w).max())


def test_embedding_quirks(golden_layers):
    G = golden_layers
    lo, hi = O.vocab_chunk_bounds(17, 1, 0)
    assert (lo, hi) == (0, 16)  # Q7: token 16 is out of chunk at tp=1
    y, ctx = O.embedding_forward(G["emb/tokens"].copy(), G["emb/E"], lo, hi, True)
    eq(y, G["emb/y"])
    assert not y[0, :3].any()
    eq(O.positional_table(32, 16), G["emb/pos"])
    eq(G["emb/pos"][:16] + y, G["emb/y_pos"])
    eq(O.output_embedding_forward(G["emb/out_x"], G["emb/E"]), G["emb/logits"])
    dx, dE = O.output_embedding_backward(G["emb/out_x"], G["emb/E"], G["emb/dlogits"])
    eq(dx, G["emb/out_dx"]); eq(dE, G["emb/dE_gemm"])
    eq(O.embedding_backward(ctx, dE.copy(), G["emb/dy"]), G["emb/dE"])
    # chunk bounds for tp>1 agree with np.array_split whenever vocab % tp == 1 (Q8)
    for v, tp in ((65, 2), (65, 4), (65, 8), (50257, 8), (17, 2)):
        sizes = [a.shape[0] for a in np.array_split(np.arange(v), tp)]
        for r in range(tp):
            lo, hi = O.vocab_chunk_bounds(v, tp, r)
            assert hi - lo == sizes[r] and lo == sum(sizes[:r])


def test_loss(golden_layers):
    G = golden_layers
    loss, grads = O.softmax_cross_entropy_world([G["loss/logits"]], G["loss/labels"].copy(), [(0, 16)])
    eq(loss, G["loss/loss"]); eq(grads[0], G["loss/dlogits"])


def test_adam(golden_layers):
    G = golden_layers
    p, m, v = G["adam/p0"], np.zeros((8, 8), np.float32), np.zeros((8, 8), np.float32)
    for i in range(3):
        p, m, v = O.adam_update(p, G[f"adam/g{i}"], m, v, 1e-3, 0.9, 0.99)
        eq(p, G[f"adam/p{i + 1}"])
    eq(m, G["adam/m"]); eq(v, G["adam/v"])
    assert p.dtype == np.float32


def _tiny_batches(G, n=3):
    return [(G[f"tiny/x{i}"], G[f"tiny/y{i}"]) for i in range(n)]


def test_tiny_transformer_bit_exact(golden_layers):
    G = golden_layers
    t = O.OracleTransformer(**TINY, rng=np.random.default_rng(42), literal=True)
    for i, (x, y) in enumerate(_tiny_batches(G)):
        tr = {}
        res = t.forward_loss(x, y)  # loss before the update
        eq(res[0][0], G[f"tiny/loss{i}"])
        t.train_step(x, y, tr)
        if i == 0:
            eq(tr["logits"][0], G["tiny/logits0"]); eq(tr["d_logits"][0], G["tiny/dlogits0"])
            eq(tr["grads"]["embedding"], G["tiny/dE0"])
            eq(tr["grads"]["block_1.qkv.weight"], G["tiny/dwqkv0_block1"])
            eq(tr["grads"]["block_1.col.weight"], G["tiny/dw1_0_block1"])
            eq(tr["grads"]["block_1.norm2.weight"], G["tiny/dn2w0_block1"])
    eq(t.ranks[0][0].params["embedding"], G["tiny/E_final"])
    eq(t.ranks[0][0].params["block_1.qkv.weight"], G["tiny/wqkv_final_block1"])


def test_tiny_transformer_fast_mode(golden_layers):
    G = golden_layers
    t = O.OracleTransformer(**TINY, rng=np.random.default_rng(42), literal=False)
    for i, (x, y) in enumerate(_tiny_batches(G)):
        loss = t.train_step(x, y)[0]
        np.testing.assert_allclose(loss, G[f"tiny/loss{i}"].mean(), rtol=2e-6)
    np.testing.assert_allclose(t.ranks[0][0].params["embedding"], G["tiny/E_final"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("tp,dp", [(2, 1), (2, 2)])
def test_tiny_world(golden_layers, tp, dp):
    """The emulated tp x dp world against the reference run as tp*dp OS processes (gloo stand-in)."""
    W = load_golden(f"tiny_tp{tp}_dp{dp}.npz")
    t = O.OracleTransformer(**TINY, rng=np.random.default_rng(42), tp=tp, dp=dp, literal=True)
    G = golden_layers
    for i, (x, y) in enumerate(_tiny_batches(G)):
        tr = {}
        res = t.forward_loss(x, y)
        for dr in range(dp):
            for trk in range(tp):
                np.testing.assert_allclose(res[dr][0], W[f"rank{dr * tp + trk}/loss{i}"], rtol=1e-6, atol=1e-6)
        t.train_step(x, y, tr)
    for dr in range(dp):
        for trk in range(tp):
            r = dr * tp + trk
            np.testing.assert_allclose(t.ranks[dr][trk].params["embedding"], W[f"rank{r}/E_final"], rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(t.ranks[dr][trk].params["block_0.col.weight"], W[f"rank{r}/w1_final_block0"],
                                       rtol=1e-5, atol=1e-7)
    if dp > 1:  # Q11: block weights diverge across DP replicas, the embedding does not
        eq(t.ranks[0][0].params["embedding"], t.ranks[1][0].params["embedding"])
        assert np.any(t.ranks[0][0].params["block_0.col.weight"] != t.ranks[1][0].params["block_0.col.weight"])


def test_reference_checkpoint_fixture_layout():
    """The checkpoint fixture written by the reference itself has the layout
    numpitron_b200.checkpoint documents: nested settings / layers / parameters dicts of float32 arrays."""
    state = np.load(GOLDEN / "tiny_checkpoint.npy", allow_pickle=True)[()]
    assert list(state) == ["model_parameters", "optimizer_parameters", "epoch", "validation_loss"]
    mp, op = state["model_parameters"], state["optimizer_parameters"]
    assert list(mp) == ["settings", "layers"]
    assert list(mp["layers"]) == ["input_embedding", "positional_encoding", "block_0", "block_1", "output_embedding"]
    w = mp["layers"]["block_0"]["layers"]["attention"]["layers"]["qkv_projection"]["parameters"]["weight"]
    assert w["data"].dtype == np.float32 and w["data"].shape == (32, 96) and w["shard_axis"] == 1 and w["gradient"] is None
    assert list(op) == ["parameters", "timestep", "learning_rate", "beta1", "beta2", "eps"]
    m = op["parameters"]["block_0"]["norm1"]["weight"]["parameters"]["momentum"]
    assert m["data"].dtype == np.float32 and m["data"].shape == (32,)


def test_oracle_loader_restatement_matches_reference_loader():
    """O.draw_batch (data/loader.py:48-54 restated) reproduces the batches the reference's own DataLoader
    yielded (tests/golden/loader_tp1.npz), from the same token stream and generator seed."""
    with np.load(GOLDEN / "loader_tp1.npz") as z:
        G = {k: z[k] for k in z.files}
    data = np.random.default_rng(1234).integers(0, 17, 5000).astype(np.uint16)   # make_golden.LOADER_TOKENS
    rng = np.random.default_rng(5)
    for i in range(6):
        x, y = O.draw_batch(data, rng, 4, 16)
        np.testing.assert_array_equal(x, G[f"x{i}"])
        np.testing.assert_array_equal(y, G[f"y{i}"])
    assert len(data) // (16 * 4) == int(G["batches_per_epoch"])


@pytest.mark.parametrize("tp", [1, 2])
def test_oracle_sampling_from_reference_checkpoint(tp):
    """The oracle, loaded from the checkpoint the reference wrote (tiny_checkpoint.npy), generates the
    reference's own token sequences (sample_tp1.npz) — greedy and multinomial at two temperatures — also
    when its vocabulary and weights are sharded over an emulated tp = 2 world."""
    state = np.load(GOLDEN / "tiny_checkpoint.npy", allow_pickle=True)[()]
    with np.load(GOLDEN / "sample_tp1.npz") as z:
        G = {k: z[k] for k in z.files}
    s = state["model_parameters"]["settings"]
    model = O.OracleTransformer(s["d_model"], s["n_heads"], s["n_layers"], s["vocab_size"], s["seq_len"], tp=tp,
                                full_params=O.full_parameters_from_checkpoint(state))
    assert O.generate(model, [3], 24, None, None) == G["greedy"].tolist()
    for name, temp in (("basic_t1", 1.0), ("basic_t40", 40.0)):
        assert O.generate(model, [3], 24, temp, np.random.default_rng(42)) == G[name].tolist(), name


def test_package_synthetic_data_matches_oracle():
    """numpitron_b200.data.synthetic (what bench.py and the multi-GPU tests draw their batches with) is the
    oracle's token stream and batch draw, which test_draw_batch_matches_reference_loader pins to the reference."""
    from numpitron_b200.data import synthetic as S

    np.testing.assert_array_equal(S.synthetic_tokens(65, 10000), O.synthetic_tokens(65, 10000))
    data = S.synthetic_tokens(17, 5000)
    for seed in (0, 99):
        a = S.draw_batch(data, np.random.default_rng(seed), 6, 16)
        b = O.draw_batch(data, np.random.default_rng(seed), 6, 16)
        np.testing.assert_array_equal(a[0], b[0])
        np.testing.assert_array_equal(a[1], b[1])
    x = np.arange(24).reshape(8, 3)
    np.testing.assert_array_equal(S.local_rows(x, 4, 2), x[4:6])


def test_workloads_and_config_vectors():
    """The BASELINE workloads partition like the reference (arange(world).reshape(pp, dp, tp)) and the committed
    oracle vectors for configs[2] / configs[3] have the layout the multi-GPU test and bench.py's parity leg read."""
    from numpitron_b200.workloads import WORKLOADS, flop_per_token, partition

    assert partition("c1", 1) == (1, 1) and partition("c1", 8) == (2, 4) and partition("c3", 8) == (8, 1)
    assert partition("c4", 8) == (2, 4) and partition("c1", 4, tp=1) == (1, 4)
    assert abs(flop_per_token(WORKLOADS["c1"]["model"]) - 70.93e6) < 0.01e6
    assert abs(flop_per_token(WORKLOADS["c3"]["model"]) - 854.4e6) < 0.1e6
    assert abs(flop_per_token(WORKLOADS["c3"]["model"], causal_skip=True) - 797.8e6) < 0.1e6
    for cfg, (tp, dp) in (("c3", (8, 1)), ("c4", (2, 4))):
        f = GOLDEN / f"config_{cfg}.npz"
        if not f.exists():
            pytest.skip(f"{f.name} not generated")
        z = np.load(f)
        m = WORKLOADS[cfg]["model"]
        assert (int(z["tp"]), int(z["dp"])) == (tp, dp)
        V_local = [len(c) for c in np.array_split(np.arange(m["vocab_size"]), tp)]
        for i in range(int(z["steps"])):
            assert z[f"x{i}"].shape == (int(z["global_batch"]), m["seq_len"]) and z[f"x{i}"].dtype == np.uint16
            for d in range(dp):
                loss = z[f"loss{i}/dp{d}"]
                assert loss.shape == (int(z["global_batch"]) // dp, m["seq_len"]) and np.all(np.isfinite(loss))
        for r in range(tp * dp):
            assert z[f"rank{r}/logits0_rowsum"].size == int(z["global_batch"]) // dp * m["seq_len"]
            assert z[f"rank{r}/dE0"].shape == z[f"rank{r}/E_final"].shape
            assert z[f"rank{r}/dE0"].shape[1] == min(int(z["e_cols"]), V_local[r % tp])


def test_reference_world_runs_the_unmodified_reference():
    """bench.py's reference arm: baseline/_ref (pip-installed from /root/reference) under oracle/mpi4py_stub, two
    OS processes tp2: its losses are the oracle's for the same world, batches and settings."""
    import bench
    from numpitron_b200.data.synthetic import draw_batch, synthetic_tokens

    if not (ROOT / "baseline" / "_ref" / "numpitron").exists():
        pytest.skip("baseline/_ref not installed (baseline/install_reference.sh)")
    from numpitron_b200 import workloads

    tiny = dict(name="tiny", model=dict(d_model=32, n_heads=4, n_layers=2, vocab_size=17, seq_len=16), seqs_per_gpu=2, sweep=None)
    workloads.WORKLOADS["tiny"] = tiny
    try:
        r = bench.run_reference_world("tiny", 2, steps=3, warmup=0, budget_s=60.0, seqs_per_replica=2)
    finally:
        workloads.WORKLOADS.pop("tiny")
    assert r["steps"] == 3 and r["tp"] == 2 and r["dp"] == 1 and r["value"] > 0
    # the same three steps on the oracle's emulated tp2 world: the reference's DataLoader draws with the model's rng
    rng = np.random.default_rng(42)
    orc = O.OracleTransformer(32, 4, 2, 17, 16, rng=rng, tp=2, dp=1, literal=True, lr=workloads.LR, beta1=workloads.BETA1,
                              beta2=workloads.BETA2)
    data = synthetic_tokens(17)
    want = [float(orc.train_step(*draw_batch(data, rng, 2, 16))[0]) for _ in range(3)]
    np.testing.assert_allclose(r["losses"], want, rtol=1e-6)
