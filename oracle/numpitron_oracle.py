"""CPU oracle for the NuMPItron training step — TEST INFRASTRUCTURE ONLY.

This file is a NumPy float32 restatement of the arithmetic of the reference's
`train_step` (`/root/reference/train_shakespeare.py:14-21`).  It is the *checker* for
`numpitron_b200`; nothing under `numpitron_b200/` may import it.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / reference legs use it.

Parity pin: `tests/test_oracle.py` checks every function here against golden vectors that
were produced by importing and running the UNMODIFIED reference in the build container
(`tests/golden/make_golden.py`, which needs `/root/reference`; the vectors themselves are
committed).  In `literal=True` mode the oracle reproduces the reference bit-for-bit on those
vectors (same NumPy expressions, same order); `literal=False` swaps two expressions for
algebraically identical, much faster ones (compact softmax backward, BLAS matmul for the
Linear contractions) — validated against the literal form in the same test file.

Ranks are emulated in-process: a "world" of tp*dp ranks is a Python list of per-rank
states that advance in lock-step; an all-reduce is a sum (or max) over that list in rank
order.  File:line citations are relative to `/root/reference/src/numpitron/`.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

F32 = np.float32


# ----------------------------------------------------------------------------------------
# initialisers — nn/core.py:10-45
# ----------------------------------------------------------------------------------------
def init_scaled(rng: np.random.Generator, shape, scale: float) -> np.ndarray:
    """"scaled_normal" is really uniform [0,1)*scale in float64, cast to float32 (core.py:28-31)."""
    return (rng.random(shape) * scale).astype(F32)


def init_for_testing(rng: np.random.Generator, shape) -> np.ndarray:
    """core.py:33-36."""
    return ((rng.random(shape) + 1) / max(shape)).astype(F32)


def positional_table(d_model: int, seq_len: int) -> np.ndarray:
    """Sinusoid table, built in float64 then stored float32 (nn/embedding.py:118-122)."""
    pos = np.arange(0, seq_len)[:, None]
    frac = np.arange(d_model, step=2) / d_model
    table = np.zeros((seq_len, d_model), dtype=F32)
    table[:, 0::2] = np.sin(pos / (10000**frac))
    table[:, 1::2] = np.cos(pos / (10000**frac))
    return table


# ----------------------------------------------------------------------------------------
# vocabulary chunk bounds — nn/embedding.py:24-37 (quirk Q7/Q8 of SURVEY.md)
# ----------------------------------------------------------------------------------------
def vocab_chunk_bounds(vocab: int, tp: int, tp_rank: int) -> tuple[int, int]:
    """[start, end) of the token ids this rank owns.

    tp == 1: (0, vocab-1) — the reference never corrects the `vocab-1` it sets in the
    constructor (embedding.py:24-25, 32-37), so token vocab-1 is treated as foreign.
    tp > 1: embedding.py:33-37 (only meaningful when vocab % tp != 0; raises otherwise).
    """
    if tp == 1:
        return 0, vocab - 1
    chunk = vocab // tp
    chunks = np.arange(0, vocab, chunk)
    chunks[1:] += vocab % tp
    return int(chunks[tp_rank]), int(chunks[tp_rank + 1])


def shard(a: np.ndarray, tp: int, tp_rank: int, axis: int | None) -> np.ndarray:
    """np.array_split sharding of a parameter (core.py:111-147)."""
    if axis is None or tp == 1:
        return a.copy()
    return np.ascontiguousarray(np.array_split(a, tp, axis=axis)[tp_rank])


# ----------------------------------------------------------------------------------------
# Linear — nn/linear.py:43-60
# ----------------------------------------------------------------------------------------
def linear_forward(x, w, b=None, literal=True):
    if literal:
        y = np.einsum("bsd, dh -> bsh", x, w)  # linear.py:44
    else:
        y = (x.reshape(-1, x.shape[-1]) @ w).reshape(*x.shape[:-1], w.shape[1])
    if b is not None:
        y = y + b  # linear.py:46-47
    return y


def linear_backward(x, w, dy, has_bias, literal=True):
    """Returns (dx, dw, db).  Gradients replace, never accumulate (linear.py:53-58)."""
    if literal:
        dw = np.einsum("bsd, bsh -> dh", x, dy)  # linear.py:55
        dx = np.einsum("bsh, dh -> bsd", dy, w)  # linear.py:60
    else:
        x2, dy2 = x.reshape(-1, x.shape[-1]), dy.reshape(-1, dy.shape[-1])
        dw = x2.T @ dy2
        dx = (dy2 @ w.T).reshape(x.shape)
    db = dy.sum(axis=(0, 1)) if has_bias else None  # linear.py:57-58
    return dx, dw, db


# ----------------------------------------------------------------------------------------
# activations — nn/activation.py
# ----------------------------------------------------------------------------------------
def relu_forward(x):
    return np.maximum(0, x)  # activation.py:38-40


def relu_backward(x, dy):
    return (x > 0) * dy  # activation.py:42-43


def softmax(x, axis=-1):
    e = np.exp(x - np.max(x, axis=axis, keepdims=True))  # activation.py:6-9
    return e / e.sum(axis=axis, keepdims=True)


def softmax_backward_jacobian(p, dp):
    """The literal (B,H,S,S,S) Jacobian form (activation.py:22-31)."""
    s = p.shape[-2]
    left = np.einsum("...ij, jk -> ...ijk", p, np.eye(s, dtype=dp.dtype))
    right = np.einsum("...ij, ...ik -> ...ijk", p, p)
    return (dp[..., None, :] @ (left - right)).reshape(p.shape)


def softmax_backward_compact(p, dp):
    """P * (dP - sum_j dP*P): algebraically identical to the Jacobian form."""
    return p * (dp - np.sum(dp * p, axis=-1, keepdims=True))


# ----------------------------------------------------------------------------------------
# Attention — nn/attention.py:34-104
# ----------------------------------------------------------------------------------------
def attention_forward(x, w_qkv, w_out, n_heads_local, literal=True):
    """x (B,S,d) -> (out (B,S,d) BEFORE the TP all-reduce, ctx).

    Columns of w_qkv are per head [q_h | k_h | v_h] (attention.py:43-45); the score scale
    is 1/sqrt(d_model) with d_model the FULL model width taken from x (attention.py:35,48).
    """
    b, s, d_model = x.shape
    qkv = linear_forward(x, w_qkv, None, literal)
    qkv = qkv.reshape(b, s, n_heads_local, -1).transpose(0, 2, 1, 3)
    q, k, v = np.split(qkv, 3, -1)
    mask = np.tri(s, s, dtype=bool)[None, None]
    scores = np.matmul(q, np.swapaxes(k, -2, -1)) / (d_model**0.5)
    scores = np.where(mask, scores, float("-inf"))
    p = softmax(scores)
    att = np.matmul(p, v).transpose(0, 2, 1, 3).reshape(b, s, -1)
    out = linear_forward(att, w_out, None, literal)
    return out, dict(x=x, att=att, q=q, k=k, v=v, p=p, mask=mask, h=n_heads_local)


def attention_backward(ctx, w_qkv, w_out, d_out, literal=True):
    """-> (dx BEFORE the TP all-reduce, dw_qkv, dw_out) — attention.py:70-104."""
    b, s, d_model = d_out.shape
    h = ctx["h"]
    d_att, dw_out, _ = linear_backward(ctx["att"], w_out, d_out, False, literal)
    d_att = d_att.reshape(b, s, h, -1).transpose(0, 2, 1, 3)
    dv = np.matmul(ctx["p"].transpose(0, 1, 3, 2), d_att)
    dp = np.matmul(d_att, ctx["v"].transpose(0, 1, 3, 2))
    ds = softmax_backward_jacobian(ctx["p"], dp) if literal else softmax_backward_compact(ctx["p"], dp)
    ds = np.where(ctx["mask"], ds, 0) / (d_model**0.5)
    dq = np.matmul(ds, ctx["k"])
    dk = np.matmul(ctx["q"].transpose(0, 1, 3, 2), ds).transpose(0, 1, 3, 2)
    dqkv = np.concatenate([dq, dk, dv], axis=-1).transpose(0, 2, 1, 3).reshape(b, s, -1)
    dx, dw_qkv, _ = linear_backward(ctx["x"], w_qkv, dqkv, False, literal)
    return dx, dw_qkv, dw_out


# ----------------------------------------------------------------------------------------
# LayerNorm — nn/normalization.py:22-52  (quirk Q12: backward divides by std WITHOUT eps)
# ----------------------------------------------------------------------------------------
def layernorm_forward(x, gamma, beta, eps=1e-5):
    mean = x.mean(axis=-1, keepdims=True)
    var = x.var(axis=-1, keepdims=True)
    y = gamma * ((x - mean) / np.sqrt(var + eps)) + beta
    return y, dict(x=x, mean=mean, var=var)


def layernorm_backward(ctx, gamma, dy, eps=1e-5):
    """-> (dx, dgamma, dbeta)."""
    x, mean, var = ctx["x"], ctx["mean"], ctx["var"]
    d_model = x.shape[-1]
    xhat = (x - mean) / np.sqrt(var + eps)
    dgamma = np.sum(dy * xhat, axis=(0, 1))
    dbeta = dy.sum(axis=(0, 1))
    wdy = gamma * dy
    c1 = np.sum(xhat * wdy, axis=-1) / d_model
    c2 = wdy.sum(axis=-1) / d_model
    dx = (wdy - c1[..., None] * xhat - c2[..., None]) / x.std(axis=-1, keepdims=True)
    return dx, dgamma, dbeta


# ----------------------------------------------------------------------------------------
# Embeddings — nn/embedding.py:44-107
# ----------------------------------------------------------------------------------------
def embedding_forward(tokens, emb, chunk_start, chunk_end, scattered=True):
    """tokens (B,S) integer, emb (d, V_local) -> (B,S,d) BEFORE the TP all-reduce."""
    mask = np.logical_or(tokens < chunk_start, tokens >= chunk_end)
    local = tokens
    if scattered:
        local = tokens - chunk_start
        local[mask] = 0
    out = np.take(emb.T, local, axis=0)
    if scattered:
        out[mask, :] = 0.0
    return out, dict(tokens=local, mask=mask, scattered=scattered)


def embedding_backward(ctx, grad, d_out):
    """np.add.at in flattened (b,s) order onto `grad` (d, V_local) — embedding.py:69-83."""
    if ctx["scattered"]:
        np.add.at(grad.T, ctx["tokens"][~ctx["mask"]], d_out[~ctx["mask"]])
    else:
        np.add.at(grad.T, ctx["tokens"], d_out)
    return grad


def output_embedding_forward(x, emb):
    return x @ emb  # embedding.py:92


def output_embedding_backward(x, emb, d_logits, literal=True):
    """-> (dx BEFORE the TP all-reduce, d_emb) — embedding.py:96-107.  The literal einsum is not a BLAS call
    (263 s per sequence at V_local = 25129, d = 1024); the fast mode contracts the same indices with matmul."""
    if literal:
        d_emb = np.einsum("bsd, bsv -> dv", x, d_logits)
    else:
        d_emb = x.reshape(-1, x.shape[-1]).T @ d_logits.reshape(-1, d_logits.shape[-1])
    dx = d_logits @ emb.T
    return dx, d_emb


# ----------------------------------------------------------------------------------------
# vocab-parallel softmax cross entropy — nn/loss.py:6-50
# ----------------------------------------------------------------------------------------
def softmax_cross_entropy_world(logits_per_rank, labels, bounds_per_rank):
    """All TP ranks at once.  logits_per_rank[r] is (B,S,V_r); labels (B,S) integer.

    Returns (loss (B,S) — identical on every rank, [d_logits_r]).
    The three all-reduces (loss.py:23-26, 34-35, 41-42) are the list reductions below.
    """
    tp = len(logits_per_rank)
    b, s, _ = logits_per_rank[0].shape
    rows = np.arange(b * s)
    flat_labels = labels.reshape(b * s)

    masks, local_labels, maxes, lg = [], [], [], []
    for r in range(tp):
        lo, hi = bounds_per_rank[r]
        m = np.logical_or(flat_labels < lo, flat_labels >= hi)
        ll = flat_labels - lo
        ll[m] = 0
        masks.append(m)
        local_labels.append(ll)
        l2 = logits_per_rank[r].reshape(b * s, -1)
        lg.append(l2)
        maxes.append(l2.max(axis=-1))
    gmax = maxes[0].copy()
    for r in range(1, tp):  # all-reduce(max)
        gmax = np.maximum(gmax, maxes[r])

    shifted, picked = [], []
    for r in range(tp):
        sh = lg[r] - gmax[:, None]
        pk = sh[rows, local_labels[r]]
        pk[masks[r]] = 0.0
        shifted.append(sh)
        picked.append(pk)
    gpicked = picked[0].copy()
    for r in range(1, tp):  # all-reduce(sum)
        gpicked = gpicked + picked[r]

    exps = [np.exp(sh) for sh in shifted]
    sums = [e.sum(axis=-1) for e in exps]
    gsum = sums[0].copy()
    for r in range(1, tp):  # all-reduce(sum)
        gsum = gsum + sums[r]

    loss = (-gpicked + np.log(gsum)).reshape(b, s)
    grads = []
    for r in range(tp):
        sm = exps[r] / gsum[:, None]
        sm[rows, local_labels[r]] -= 1 - masks[r]
        grads.append(sm.reshape(b, s, -1))
    return loss, grads


# ----------------------------------------------------------------------------------------
# Adam — optimizer/adam.py:15-45 (quirk Q10: timestep stays 1, eps outside the sqrt)
# ----------------------------------------------------------------------------------------
def adam_update(p, g, m, v, lr, beta1, beta2, eps=1e-7, timestep=1):
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * np.power(g, 2)
    mh = m / (1 - beta1**timestep)
    vh = v / (1 - beta2**timestep)
    p = p - (lr * mh) / (np.sqrt(vh) + eps)
    return p, m, v


# ----------------------------------------------------------------------------------------
# model construction in the reference's RNG order
# ----------------------------------------------------------------------------------------
PARAM_SHARD_AXIS = {
    "embedding": 1,
    "norm1.weight": None, "norm1.bias": None,
    "qkv.weight": 1, "out.weight": 0,
    "norm2.weight": None, "norm2.bias": None,
    "col.weight": 1, "col.bias": 0,
    "row.weight": 0, "row.bias": None,
}


def build_full_parameters(d_model, n_heads, n_layers, vocab, rng) -> dict:
    """Full (unsharded) parameter dict keyed 'embedding' / 'block_{i}.<name>'.

    Draw order: E, then per block norm1.weight, qkv, out, norm2.weight, col, row
    (transformer.py:19-46, transformer_block.py:12-17, attention.py:12-27, mlp.py:19-33).
    LayerNorm weight scale 0.2, Linear scale 0.006, embedding scale 1.0; biases zero.
    """
    P = {"embedding": init_scaled(rng, (d_model, vocab), 1.0)}
    for i in range(n_layers):
        pre = f"block_{i}."
        P[pre + "norm1.weight"] = init_scaled(rng, (d_model,), 0.2)
        P[pre + "norm1.bias"] = np.zeros((d_model,), F32)
        P[pre + "qkv.weight"] = init_scaled(rng, (d_model, 3 * d_model), 0.006)
        P[pre + "out.weight"] = init_scaled(rng, (d_model, d_model), 0.006)
        P[pre + "norm2.weight"] = init_scaled(rng, (d_model,), 0.2)
        P[pre + "norm2.bias"] = np.zeros((d_model,), F32)
        P[pre + "col.weight"] = init_scaled(rng, (d_model, 4 * d_model), 0.006)
        P[pre + "col.bias"] = np.zeros((4 * d_model,), F32)
        P[pre + "row.weight"] = init_scaled(rng, (4 * d_model, d_model), 0.006)
        P[pre + "row.bias"] = np.zeros((d_model,), F32)
    return P


def shard_axis_of(name: str):
    return PARAM_SHARD_AXIS[name.split(".", 1)[1] if name.startswith("block_") else name]


@dataclass
class RankState:
    """Parameters + Adam state of one (dp_rank, tp_rank)."""

    tp_rank: int
    dp_rank: int
    params: dict
    m: dict = field(default_factory=dict)
    v: dict = field(default_factory=dict)
    grads: dict = field(default_factory=dict)


class OracleTransformer:
    """The reference Transformer + Adam on an emulated tp x dp world.

    rank = dp_rank * tp + tp_rank (distributed/__init__.py:149-162).  Every rank is built
    from the same full parameters (every reference rank draws the full model from an
    identically seeded rng, train_shakespeare.py:50-61) and then sharded.
    """

    def __init__(self, d_model, n_heads, n_layers, vocab, seq_len, rng=None, tp=1, dp=1,
                 full_params=None, literal=True, lr=1e-4, beta1=0.9, beta2=0.99, eps=1e-7):
        self.d, self.H, self.L, self.V, self.S = d_model, n_heads, n_layers, vocab, seq_len
        self.tp, self.dp, self.literal = tp, dp, literal
        self.lr, self.beta1, self.beta2, self.eps = lr, beta1, beta2, eps
        self.full = full_params if full_params is not None else build_full_parameters(
            d_model, n_heads, n_layers, vocab, rng)
        self.pos = positional_table(d_model, seq_len)
        self.h_local = n_heads // tp  # attention.py:37-41 — no divisibility check (Q5)
        # test hooks: with keep=True every train_step leaves the per-rank logits (of the forward) and the
        # per-rank gradients (as they reach Adam) in last_logits[dp_rank][tp_rank] / last_grads[...]
        self.keep = False
        self.last_logits, self.last_grads = [], []
        self.bounds = [vocab_chunk_bounds(vocab, tp, r) for r in range(tp)]
        self.ranks: list[list[RankState]] = []
        for dr in range(dp):
            row = []
            for tr in range(tp):
                ps = {k: shard(v, tp, tr, shard_axis_of(k)) for k, v in self.full.items()}
                st = RankState(tr, dr, ps)
                st.m = {k: np.zeros_like(v) for k, v in ps.items()}
                st.v = {k: np.zeros_like(v) for k, v in ps.items()}
                row.append(st)
            self.ranks.append(row)

    # -- helpers -----------------------------------------------------------------------
    @staticmethod
    def _allreduce(xs):
        tot = xs[0].copy()
        for x in xs[1:]:
            tot = tot + x
        return [tot.copy() for _ in xs] if len(xs) > 1 else [tot]

    def split_batch(self, a):
        """loader.py:56-80: Scatterv along batch == array_split(batch, dp)."""
        return [np.ascontiguousarray(c) for c in np.array_split(a, self.dp, axis=0)]

    # -- one DP replica (all its TP ranks in lock-step) -------------------------------------
    def _forward_replica(self, row, tokens, trace=None):
        tp, lit = self.tp, self.literal
        ctx = {"emb": [], "blocks": []}
        parts = []
        for st in row:
            lo, hi = self.bounds[st.tp_rank]
            o, c = embedding_forward(tokens.copy(), st.params["embedding"], lo, hi, True)
            parts.append(o)
            ctx["emb"].append(c)
        x = self._allreduce(parts) if tp > 1 else parts
        x = [self.pos[: tokens.shape[1], :] + xi for xi in x]  # embedding.py:125-128
        if trace is not None:
            trace["embedded"] = x[0].copy()
        for i in range(self.L):
            pre = f"block_{i}."
            bctx = {"att": [], "ln1": [], "mlp": [], "ln2": []}
            outs = []
            for st, xi in zip(row, x):
                o, c = attention_forward(xi, st.params[pre + "qkv.weight"],
                                         st.params[pre + "out.weight"], self.h_local, lit)
                outs.append(o)
                bctx["att"].append(c)
            outs = self._allreduce(outs) if tp > 1 else outs
            hs = []
            for st, xi, oi in zip(row, x, outs):
                y, c = layernorm_forward(oi, st.params[pre + "norm1.weight"], st.params[pre + "norm1.bias"])
                hs.append(y + xi)  # transformer_block.py:20 (Q1)
                bctx["ln1"].append(c)
            outs = []
            for st, hi_ in zip(row, hs):
                a = linear_forward(hi_, st.params[pre + "col.weight"], st.params[pre + "col.bias"], lit)
                r = relu_forward(a)
                bias = st.params[pre + "row.bias"] if st.tp_rank == 0 else None  # mlp.py:35 (Q17)
                o = linear_forward(r, st.params[pre + "row.weight"], bias, lit)
                outs.append(o)
                bctx["mlp"].append(dict(x=hi_, a=a, r=r))
            outs = self._allreduce(outs) if tp > 1 else outs
            x = []
            for st, hi_, oi in zip(row, hs, outs):
                y, c = layernorm_forward(oi, st.params[pre + "norm2.weight"], st.params[pre + "norm2.bias"])
                x.append(y + hi_)  # transformer_block.py:21
                bctx["ln2"].append(c)
            ctx["blocks"].append(bctx)
            if trace is not None:
                trace[f"block_{i}"] = x[0].copy()
        ctx["final"] = x
        logits = [output_embedding_forward(xi, st.params["embedding"]) for st, xi in zip(row, x)]
        return logits, ctx

    def _backward_replica(self, row, ctx, d_logits, trace=None):
        tp, lit = self.tp, self.literal
        dxs = []
        for st, xi, dl in zip(row, ctx["final"], d_logits):
            dx, d_emb = output_embedding_backward(xi, st.params["embedding"], dl, lit)
            st.grads["embedding"] = d_emb
            dxs.append(dx)
        d = self._allreduce(dxs) if tp > 1 else dxs
        if trace is not None:
            trace["d_final"] = d[0].copy()
        for i in reversed(range(self.L)):
            pre = f"block_{i}."
            bctx = ctx["blocks"][i]
            # transformer_block.py:24-27 — the residual gradient is dropped (Q2)
            nxt = []
            for k, (st, di) in enumerate(zip(row, d)):
                dn, dg, db = layernorm_backward(bctx["ln2"][k], st.params[pre + "norm2.weight"], di)
                st.grads[pre + "norm2.weight"], st.grads[pre + "norm2.bias"] = dg, db
                mc = bctx["mlp"][k]
                has_bias = st.tp_rank == 0
                dr, dw_row, db_row = linear_backward(mc["r"], st.params[pre + "row.weight"], dn, has_bias, lit)
                st.grads[pre + "row.weight"] = dw_row
                if has_bias:
                    st.grads[pre + "row.bias"] = db_row
                da = relu_backward(mc["a"], dr)
                dxm, dw_col, db_col = linear_backward(mc["x"], st.params[pre + "col.weight"], da, True, lit)
                st.grads[pre + "col.weight"], st.grads[pre + "col.bias"] = dw_col, db_col
                nxt.append(dxm)
            d = self._allreduce(nxt) if tp > 1 else nxt
            nxt = []
            for k, (st, di) in enumerate(zip(row, d)):
                dn, dg, db = layernorm_backward(bctx["ln1"][k], st.params[pre + "norm1.weight"], di)
                st.grads[pre + "norm1.weight"], st.grads[pre + "norm1.bias"] = dg, db
                dxa, dw_qkv, dw_out = attention_backward(bctx["att"][k], st.params[pre + "qkv.weight"],
                                                         st.params[pre + "out.weight"], dn, lit)
                st.grads[pre + "qkv.weight"], st.grads[pre + "out.weight"] = dw_qkv, dw_out
                nxt.append(dxa)
            d = self._allreduce(nxt) if tp > 1 else nxt
            if trace is not None:
                trace[f"d_block_{i}"] = d[0].copy()
        for k, (st, di) in enumerate(zip(row, d)):
            st.grads["embedding"] = embedding_backward(ctx["emb"][k], st.grads["embedding"], di)

    # -- public -------------------------------------------------------------------------------
    def forward_loss(self, x, y, trace=None):
        """Forward + loss on the whole emulated world.  Returns per-replica (loss, ctx, dlogits)."""
        xs, ys = self.split_batch(x), self.split_batch(y)
        out = []
        self.last_logits = []
        for row, xb, yb in zip(self.ranks, xs, ys):
            logits, ctx = self._forward_replica(row, xb, trace)
            if self.keep:
                self.last_logits.append(logits)
            if trace is not None:
                trace["logits"] = [l.copy() for l in logits]
            loss, dl = softmax_cross_entropy_world(logits, yb.copy(), self.bounds)
            out.append((loss, ctx, dl))
        return out

    def train_step(self, x, y, trace=None):
        """train_shakespeare.py:14-21.  Returns the per-DP-replica loss.mean() list."""
        res = self.forward_loss(x, y, trace)
        for row, (loss, ctx, dl) in zip(self.ranks, res):
            if trace is not None:
                trace["d_logits"] = [g.copy() for g in dl]
            self._backward_replica(row, ctx, dl, trace)
        if self.dp > 1:  # transformer.py:52-57 — ONLY the embedding gradient (Q11), SUM
            for tr in range(self.tp):
                tot = self.ranks[0][tr].grads["embedding"].copy()
                for dr in range(1, self.dp):
                    tot = tot + self.ranks[dr][tr].grads["embedding"]
                for dr in range(self.dp):
                    self.ranks[dr][tr].grads["embedding"] = tot.copy()
        if trace is not None:
            trace["grads"] = {k: v.copy() for k, v in self.ranks[0][0].grads.items()}
        if self.keep:
            self.last_grads = [[dict(st.grads) for st in row] for row in self.ranks]
        for row in self.ranks:
            for st in row:
                for name, g in st.grads.items():
                    st.params[name], st.m[name], st.v[name] = adam_update(
                        st.params[name], g, st.m[name], st.v[name],
                        self.lr, self.beta1, self.beta2, self.eps)
                st.grads = {}
        return [loss.mean() for (loss, _, _) in res]


def synthetic_tokens(vocab: int, n: int = 2_000_000, seed: int = 1234) -> np.ndarray:
    """The synthetic uint16 token stream named in SURVEY.md §8(d)."""
    return np.random.default_rng(seed).integers(0, vocab, n).astype(np.uint16)


def draw_batch(data: np.ndarray, rng: np.random.Generator, batch: int, seq_len: int):
    """One global batch exactly as data/loader.py:48-54 draws it (labels = next token)."""
    start = rng.integers(0, len(data) - 1 - seq_len, size=batch)[:, None]
    rng_ = np.arange(seq_len)
    return data[start + rng_], data[start + 1 + rng_]


# ---------------------------------------------------------------------------------------------------
# checkpoint layout, sampling (test infrastructure like everything in this file)
# ---------------------------------------------------------------------------------------------------
def full_parameters_from_checkpoint(state: dict) -> dict:
    """Map the reference's checkpoint tree (train_shakespeare.py:30-41: model.to_dict() of the gathered
    model, nn/model.py:42-44, nn/core.py:96-99) onto the flat keys of `build_full_parameters`."""
    layers = state["model_parameters"]["layers"]
    P = {"embedding": layers["input_embedding"]["parameters"]["embedding"]["data"]}
    i = 0
    while f"block_{i}" in layers:
        blk, pre = layers[f"block_{i}"]["layers"], f"block_{i}."
        att, mlp = blk["attention"]["layers"], blk["mlp"]["layers"]
        P[pre + "norm1.weight"] = blk["norm1"]["parameters"]["weight"]["data"]
        P[pre + "norm1.bias"] = blk["norm1"]["parameters"]["bias"]["data"]
        P[pre + "qkv.weight"] = att["qkv_projection"]["parameters"]["weight"]["data"]
        P[pre + "out.weight"] = att["out_projection"]["parameters"]["weight"]["data"]
        P[pre + "norm2.weight"] = blk["norm2"]["parameters"]["weight"]["data"]
        P[pre + "norm2.bias"] = blk["norm2"]["parameters"]["bias"]["data"]
        P[pre + "col.weight"] = mlp["column_linear"]["parameters"]["weight"]["data"]
        P[pre + "col.bias"] = mlp["column_linear"]["parameters"]["bias"]["data"]
        P[pre + "row.weight"] = mlp["row_linear"]["parameters"]["weight"]["data"]
        P[pre + "row.bias"] = mlp["row_linear"]["parameters"]["bias"]["data"]
        i += 1
    return P


def basic_sample(logits_last: np.ndarray, temperature: float, rng: np.random.Generator) -> int:
    """sample_shakespeare.py:13-25 on the gathered last-position logits: softmax of
    (logits / temperature) in float64, one multinomial draw, arg-max of the one-hot."""
    probabilities = softmax((logits_last.astype(F32) / temperature).astype(np.float64))
    return int(rng.multinomial(1, probabilities).argmax())


def greedy_sample(logit_shards) -> int:
    """sample_shakespeare.py:28-49 over the per-rank vocabulary shards of the last position: every rank
    reports (max value, arg-max + tp_rank * len(its shard)); the index of the rank with the largest value
    wins.  The offset uses the rank's OWN shard length (:37), so with an uneven vocabulary split
    (17 -> 9 + 8) an arg-max in a later shard comes out shifted — reproduced, like the other quirks."""
    if isinstance(logit_shards, np.ndarray):
        logit_shards = [logit_shards]
    values = [float(s.max()) for s in logit_shards]
    indices = [int(s.argmax()) + r * len(s) for r, s in enumerate(logit_shards)]
    return int(np.float32(indices[int(np.argmax(values))]))


def generate(model: "OracleTransformer", tokens, num_samples: int, temperature, rng) -> list:
    """The loop of sample_shakespeare.py:67-75 (sliding context window, full forward per token)."""
    tokens, out = list(tokens), []
    for _ in range(num_samples):
        tokens = tokens[-model.S:]
        logits, _ = model._forward_replica(model.ranks[0], np.asarray([tokens], dtype=np.uint16))
        shards = [l[0, -1, :] for l in logits]
        last = np.concatenate(shards)                          # all_gather over the TP group
        nxt = greedy_sample(shards) if temperature is None else basic_sample(last, temperature, rng)
        tokens.append(nxt)
        out.append(nxt)
    return out
