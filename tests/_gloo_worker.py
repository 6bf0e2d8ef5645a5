"""One rank of the world-size-N gloo tests (launched by tests/test_distributed_gloo.py).
Checks the HOST-side logic of the tp x dp world: group layout, all_reduce / scatter / all_gather
semantics, parameter sharding, vocab chunk bounds, DP batch split.  No kernels are launched."""
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from numpitron_b200 import distributed as npdist  # noqa: E402
from numpitron_b200.nn import Transformer  # noqa: E402
from numpitron_b200.optimizer import Adam  # noqa: E402
from oracle import numpitron_oracle as O  # noqa: E402


def main(tp, dp):
    world, rank = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"])
    npdist.init(tp_size=world)          # what importing the reference's transformer module does (Q15)
    npdist.init(tp_size=tp, dp_size=dp)  # re-init with the real sizes must be tolerated
    assert npdist.world_size() == world and npdist.world_rank() == rank
    # rank = dp_rank * tp + tp_rank  (arange(world).reshape(pp, dp, tp))
    assert npdist.tensor_parallel_rank() == rank % tp and npdist.data_parallel_rank() == rank // tp
    assert npdist.tensor_parallel_group().ranks == tuple(range(rank // tp * tp, rank // tp * tp + tp))
    assert npdist.data_parallel_group().ranks == tuple(range(rank % tp, world, tp))
    assert npdist.pipeline_parallel_size() == 1

    # all_reduce: in place, returns None, sum and max, numpy and torch buffers
    a = np.full((3, 2), float(rank + 1), np.float32)
    assert npdist.all_reduce(a, group=npdist.tensor_parallel_group()) is None
    tp_ranks = npdist.tensor_parallel_group().ranks
    np.testing.assert_array_equal(a, np.full((3, 2), sum(r + 1 for r in tp_ranks), np.float32))
    t = torch.full((4,), float(rank))
    npdist.all_reduce(t, op="max", group=npdist.data_parallel_group())
    assert float(t[0]) == max(npdist.data_parallel_group().ranks)
    w = np.array([1.0], np.float32)
    npdist.all_reduce(w)
    assert w[0] == world

    # model sharding == np.array_split of the full parameters, chunk bounds as the reference sets them
    cfg = dict(d_model=32, n_heads=4, n_layers=2, vocab_size=17, seq_len=16)
    model = Transformer(**cfg, rng=np.random.default_rng(42))
    opt = Adam(model, 1e-4, 0.9, 0.99)
    full = O.build_full_parameters(32, 4, 2, 17, np.random.default_rng(42))
    if rank != 0:  # scatter takes rank-0-of-the-TP-group's values: poison the others to prove it
        model.block_0.mlp.column_linear.weight.data.add_(rank % tp * 100.0)
    model.scatter()
    opt.scatter()
    tr = npdist.tensor_parallel_rank()
    names = {"block_0.qkv.weight": model.block_0.attention.qkv_projection.weight,
             "block_0.out.weight": model.block_0.attention.out_projection.weight,
             "block_0.col.weight": model.block_0.mlp.column_linear.weight,
             "block_0.col.bias": model.block_0.mlp.column_linear.bias,
             "block_1.row.weight": model.block_1.mlp.row_linear.weight,
             "block_1.norm1.weight": model.block_1.norm1.weight,
             "embedding": model.input_embedding.embedding}
    for name, p in names.items():
        want = O.shard(full[name], tp, tr, O.shard_axis_of(name))
        np.testing.assert_array_equal(p.data.numpy(), want, err_msg=name)
    lo, hi = O.vocab_chunk_bounds(17, tp, tr)
    assert (npdist.get_var("embedding_chunk_start"), npdist.get_var("embedding_chunk_end")) == (lo, hi)
    if tp > 1:
        # bookkeeping of the small-vocabulary gather path (InputEmbedding.gathers_full_table): every rank knows every
        # rank's chunk bounds and shard width; 17 % 2 == 1, so the reference's bounds coincide with the shards
        emb = model.input_embedding
        assert emb._chunks == [O.vocab_chunk_bounds(17, tp, r)[0] for r in range(tp)] + [O.vocab_chunk_bounds(17, tp, tp - 1)[1]]
        assert emb._widths == [O.shard(full["embedding"], tp, r, 1).shape[1] for r in range(tp)]
        assert sum(emb._widths) == 17 and emb._chunks[tr + 1] - emb._chunks[tr] == emb.embedding.data.shape[1]
        assert emb.gathers_full_table(10_000) is False   # no CUDA device here: the path is GPU-only
    assert model.block_0.mlp.row_linear.use_bias == (tr == 0)                      # Q17
    assert model.is_scattered and model.block_1.attention.is_scattered
    m = opt.parameters["block_0"]["mlp"]["column_linear"]["weight"].momentum.data
    assert tuple(m.shape) == tuple(model.block_0.mlp.column_linear.weight.data.shape)
    assert model.block_0.attention._local_heads() == 4 // tp                         # Q5

    # all_gather restores the full parameters (checkpoint path) and resets the chunk bounds
    model.all_gather()
    np.testing.assert_array_equal(model.block_0.attention.qkv_projection.weight.data.numpy(), full["block_0.qkv.weight"])
    np.testing.assert_array_equal(model.input_embedding.embedding.data.numpy(), full["embedding"])
    assert npdist.get_var("embedding_chunk_end") == 16

    # DP batch split: every rank draws the same global batch, takes array_split(...)[dp_rank]
    data = O.synthetic_tokens(17, 4096)
    x, _ = O.draw_batch(data, np.random.default_rng(99), 8, 16)
    mine = np.array_split(x, dp, axis=0)[npdist.data_parallel_rank()]
    gathered = np.zeros((dp,) + mine.shape, np.float32)
    gathered[npdist.data_parallel_rank()] = mine
    npdist.all_reduce(gathered, group=npdist.data_parallel_group())
    np.testing.assert_array_equal(gathered.reshape(x.shape), x.astype(np.float32))

    # samplers across the tensor-parallel group (sample_shakespeare.py:13-49): each rank holds its shard of
    # the last position's logits; the temperature sampler all-gathers them, the greedy one all-reduces
    # (value, index) pairs — including the reference's index offset by the rank's OWN shard length
    from numpitron_b200.sample import basic_sample, greedy_sample

    npdist.init(tp_size=tp, dp_size=dp)
    for seed in range(6):
        full_logits = (np.random.default_rng(100 + seed).standard_normal(17) * 3).astype(np.float32)
        shards = np.array_split(full_logits, tp)
        local = shards[tr][None, None, :]
        assert basic_sample(local, 0.7, np.random.default_rng(11), 17) == O.basic_sample(full_logits, 0.7, np.random.default_rng(11))
        assert greedy_sample(local) == O.greedy_sample(shards)

    # checkpoint round trip under tensor parallelism (train_shakespeare.py:30-41,87-93): gather, write the
    # full tree, scatter again; the file holds the FULL parameters and Adam moments on every rank
    import tempfile

    from numpitron_b200.checkpoint import load_checkpoint, save_checkpoint

    model.scatter()
    with tempfile.TemporaryDirectory() as tmp:
        path = Path(tmp) / f"ckpt_rank{rank}.npy"
        save_checkpoint(model, opt, 3, 1.25, path)
        assert model.is_scattered and tuple(model.block_0.attention.qkv_projection.weight.data.shape) == (32, 96 // tp)
        state = np.load(path, allow_pickle=True)[()]
        assert state["epoch"] == 3 and state["validation_loss"] == 1.25
        blk = state["model_parameters"]["layers"]["block_0"]["layers"]
        np.testing.assert_array_equal(blk["attention"]["layers"]["qkv_projection"]["parameters"]["weight"]["data"],
                                      full["block_0.qkv.weight"])
        np.testing.assert_array_equal(blk["mlp"]["layers"]["row_linear"]["parameters"]["weight"]["data"],
                                      full["block_0.row.weight"])
        np.testing.assert_array_equal(O.full_parameters_from_checkpoint(state)["embedding"], full["embedding"])
        mom = state["optimizer_parameters"]["parameters"]["block_0"]["mlp"]["column_linear"]["weight"]["parameters"]["momentum"]
        assert mom["data"].shape == (32, 128) and mom["shard_axis"] == 1
        t2, o2, epoch, vloss = load_checkpoint(path)
        assert epoch == 3 and vloss == 1.25 and not t2.is_scattered
        np.testing.assert_array_equal(t2.block_1.mlp.column_linear.weight.data.numpy(), full["block_1.col.weight"])
        t2.scatter()
        o2.scatter()
        np.testing.assert_array_equal(t2.block_1.mlp.column_linear.weight.data.numpy(),
                                      O.shard(full["block_1.col.weight"], tp, tr, 1))
    print(f"rank {rank} ok", flush=True)


if __name__ == "__main__":
    main(int(sys.argv[1]), int(sys.argv[2]))
