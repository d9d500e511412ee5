"""``shard_model``: turn a replicated-initialised model into its row-sharded form, in place."""
from __future__ import annotations

import torch.nn as nn

from .sharded import ShardedTables, reduce_dense_grads


def shard_model(model, pg=None, device=None, backend=None, transport=None, mode=None, dedup=None, init_seed=None, hybrid=None,
                replicate_max_rows: int = 1 << 17, hot_rows: int = 0):
    """Every rank calls this with an identically initialised ``model`` (same seed).  The model's tables
    (and, for DeepFM, its first-order tables, which share the ids) are cut into this rank's rows, fused
    into one shard per width on ``device`` and the full tables are dropped; the dense part stays
    replicated.  Afterwards use::

        loss = model.training_step(batch, i)
        (loss / world).backward()            # tables are updated inside backward
        model.reduce_dense_grads()           # one all-reduce for the tower
        optimizer.step()

    ``hybrid`` (peer mode; default: on when the model qualifies and has small tables or DeepFM's fused terms): tables with at
    most ``replicate_max_rows`` rows stay replicated on every rank (their gradients are all-reduced densely), the others are
    row-sharded -- ``parallel.hybrid.HybridShardedTables``.  ``hot_rows`` > 0 additionally replicates the first ``hot_rows``
    rows of every large direct-id table (the hot rows of a frequency-ordered vocabulary); only the tails are sharded.
    """
    groups = model._groups
    full = [[g.tables[n] for n in g.names] for g in groups]
    if mode is None:
        mode = "peer" if (backend is None and device is not None and str(device).startswith("cuda")) or transport is not None else "a2a"
    if mode == "peer":
        # B200 path: rows are read from / gradients pulled through NVLink peer mappings inside the kernels
        from .peer import IpcTransport, PeerShardedTables
        if transport is None:
            transport = IpcTransport(pg, device)
        from .hybrid import HybridShardedTables, hybrid_eligible
        twins = full[1] if len(full) == 2 and all(t.embedding_dim == 1 for t in full[1]) else None
        can = len(full) <= 2 and (len(full) == 1 or twins is not None) and hybrid_eligible(full[0], twins)
        if hybrid is None:                    # (an explicit dedup=True asks for the fully sharded, de-duplicated exchange)
            hybrid = can and dedup is not True and (twins is not None or any(t.num_embeddings <= replicate_max_rows for t in full[0]))
        if dedup is None:                     # fully sharded: the requester-side sort pays off once nearly all rows are remote:
            dedup = transport.world > 4       # measured -10 % at 2 GPUs, +6 % at 8
        if hybrid and not can:
            raise NotImplementedError("hybrid placement needs single-id sum-pooled tables of one width (16 / 32 / 64)")
        if hybrid:
            sharded = HybridShardedTables(groups[0].names, full[0], twins, transport, device, fm=bool(getattr(model, "_fm_term", False)),
                                          replicate_max_rows=replicate_max_rows, init_seed=init_seed, hot_rows=hot_rows)
        else:
            # tables declared on the meta device (model built with table_device='meta') are created shard by shard, in place
            sharded = PeerShardedTables(groups[0].names, full, transport, device, dedup=dedup, init_seed=init_seed)
    elif device is not None:
        # shards are built on the target device straight from the (host) full tables
        import torch
        with torch.device(device):
            sharded = ShardedTables(groups[0].names, full, pg, backend)
    else:
        sharded = ShardedTables(groups[0].names, full, pg, backend)
    model._sharded = sharded
    for g in groups:                       # drop the replicated tables
        for n in list(g.tables.keys()):
            del g.tables[n]
    model._pg = pg
    def _reduce():
        if getattr(model, "_dense_grads_reduced_in_backward", False):
            return                              # hybrid placement: already summed with the replicated tables' gradients
        flat = getattr(model, "_flat_dense_grad", None)
        if flat is not None:                    # one all-reduce over the flat gradient buffer, nothing to copy
            import torch.distributed as dist
            dist.all_reduce(flat, group=pg)
        else:
            reduce_dense_grads(model.dense_parameters(), pg)
    model.reduce_dense_grads = _reduce

    table_prefix = []
    for g in groups:                       # 'embeddings', 'linear_embeddings': the ModuleDicts the tables lived in
        table_prefix.append(next((n for n, m in model.named_children() if m is g.tables), None))

    def full_state_dict():
        """Collective.  The state dict of the UNSHARDED model -- every table gathered back to ``[V, D]`` under its
        original key -- so that ``Trainer.save_ckpt`` output (``torchctr/trainer.py:353-496``) written from it loads
        into the reference ``DNN`` or into an unsharded model of this package."""
        if not hasattr(sharded, "export_full_tables"):
            raise NotImplementedError("full_state_dict(): the all-to-all formulation (ShardedTables) cannot gather its tables; "
                                      "use the peer-memory mode (mode='peer') for checkpoints in the reference's format")
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items() if not k.startswith("_sharded.")}
        for w, prefix in enumerate(table_prefix):
            if prefix is None:
                continue
            for name, t in zip(sharded.names, sharded.export_full_tables(w)):
                sd[f"{prefix}.{name}.weight"] = t
        return sd

    def full_table_optimizer_state():
        """Collective.  The fused-update state of the sharded tables gathered to full ``[V, D]`` tensors, keyed like
        ``CTRModelBase.table_optimizer_state_dict`` ({'embeddings.<name>': {'state0', 'state1'}}): together with
        ``full_state_dict()`` a complete, sharding-independent checkpoint."""
        out = {}
        for w, prefix in enumerate(table_prefix):
            if prefix is None:
                continue
            s0, s1 = sharded.export_full_optimizer_state(w)
            if s0 is None:
                continue
            for f, name in enumerate(sharded.names):
                out[f"{prefix}.{name}"] = {"state0": s0[f], "state1": None if s1 is None else s1[f]}
        return out

    def load_full_state(state_dict, table_optimizer_state=None):
        """Collective inverse: every rank passes the same full state dict (reference format) and optional optimizer state."""
        import torch
        dense = {k: v for k, v in state_dict.items() if not any(k.startswith(p + ".") for p in table_prefix if p)}
        own = model.state_dict()
        model.load_state_dict({**{k: v for k, v in own.items() if k.startswith("_sharded.")}, **dense}, strict=False)
        for w, prefix in enumerate(table_prefix):
            if prefix is None:
                continue
            sharded.load_full_tables([state_dict[f"{prefix}.{n}.weight"] for n in sharded.names], w)
            if table_optimizer_state:
                ent = [table_optimizer_state.get(f"{prefix}.{n}") for n in sharded.names]
                if all(e is not None for e in ent):
                    s1 = None if ent[0]["state1"] is None else [e["state1"] for e in ent]
                    sharded.load_full_optimizer_state(([e["state0"] for e in ent], s1), w)
    model.full_table_optimizer_state = full_table_optimizer_state
    model.load_full_state = load_full_state
    model.full_state_dict = full_state_dict
    return model
