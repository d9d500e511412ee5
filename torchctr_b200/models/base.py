"""Shared model plumbing: tables from ``feat_configs``, fused lookup, reference-shaped tower.

Mirrors the construction rules of ``torchctr/models/dnn.py:10-46`` (same attribute names, same
module order, hence the same ``state_dict`` keys) and its step protocol (``:72-82``).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..nn.embedding import EmbeddingTable, PlanLink, PooledLookupGroup, take_blocked_offer
from ..nn.linear import linear_tc
from ..nn.head import head_eligible, logit_bce
from ..nn import tower as _tower
from ..nn.linear import matmul_precision
from ..nn.tower import block_is_fusable, prepare_first_block, tower_block
from ..nn.vocab import VocabIndex


def split_feature_configs(feat_configs):
    """(sparse configs, dense input width) -- dnn.py:17-29 incl. its error messages."""
    sparse, dense_width = [], 0
    for cfg in feat_configs:
        kind = cfg["type"]
        if kind == "sparse":
            if "emb_dim" not in cfg:
                raise ValueError("emb_dim must be specified for sparse features.")
            sparse.append(cfg)
        elif kind == "dense" and cfg.get("islist"):
            dense_width += 3      # dnn.py:24-25: 3 dense inputs per dense sequence feature
        elif kind == "dense":
            dense_width += 1
        else:
            raise ValueError(f'Unsupported feature type: {cfg["type"]}')
    return sparse, dense_width


def make_tower(input_dim: int, hidden_units, p_drop: float = 0.5) -> nn.Sequential:
    """[Linear, BatchNorm1d, ReLU, Dropout(0.5)] * k + Linear(., 1) -- dnn.py:35-46."""
    layers, width = [], input_dim
    for h in hidden_units:
        layers += [nn.Linear(width, h), nn.BatchNorm1d(h), nn.ReLU(), nn.Dropout(p=p_drop)]
        width = h
    layers.append(nn.Linear(width, 1))
    return nn.Sequential(*layers)


def table_from_config(cfg, device=None) -> EmbeddingTable:
    """Table for one sparse feature.  ``raw_ids: True`` (extension) means the batch carries raw
    category ids rather than rows: with ``hash_buckets`` the murmur3 bucket of
    ``torchctr/transformer.py:487-490`` is computed inside the lookup, without it the ids go
    through a device vocabulary that grows while training (``transformer.py:451-498``)."""
    kind, vocab = "direct", None
    if cfg.get("raw_ids"):
        if cfg.get("hash_buckets"):
            kind = "hash"
        else:
            kind = "vocab"
            vocab = VocabIndex(capacity=cfg.get("vocab_capacity", 1 << 16), min_freq=cfg.get("min_freq", 0))
    kw = {} if device is None else {"device": device}
    t = _make_table(cfg, kind, vocab, kw)
    if kind == "vocab":
        t.vocab_max_rows = int(cfg.get("vocab_max_rows", 1 << 20))      # row capacity of the table when it is row-sharded
    return t


def _make_table(cfg, kind, vocab, kw) -> EmbeddingTable:
    return EmbeddingTable(cfg["num_embeddings"], cfg["emb_dim"], pooling=cfg.get("pooling", "sum"), index_kind=kind,
                          hash_seed=cfg.get("seed", 0), vocab=vocab, use_id_weight=bool(cfg.get("use_weight", False)), **kw)


class CTRModelBase(nn.Module):
    """Holds ``feat_configs``, ``embeddings`` (ModuleDict of tables keyed by feature name) and the
    fused lookup; subclasses add the interaction and the tower."""

    def __init__(self, feat_configs, table_device=None):
        """``table_device='meta'``: the tables are declared but not allocated -- for models whose tables only ever exist as
        row shards (``parallel.shard_model`` fills every rank's rows in place, BASELINE config 4: 2^30 rows x 64) or that are
        materialised on the device with ``materialize_tables``."""
        super().__init__()
        self.feat_configs = feat_configs
        self._sparse, self._dense_width = split_feature_configs(feat_configs)
        self._table_device = table_device
        self.embeddings = nn.ModuleDict({c["name"]: table_from_config(c, table_device) for c in self._sparse})
        self._names = [c["name"] for c in self._sparse]
        self._sparse_width = sum(c["emb_dim"] for c in self._sparse)
        self._lookup = PooledLookupGroup(self._names, self.embeddings)
        self._groups = [self._lookup]
        self._sharded = None              # set by torchctr_b200.parallel.shard_model

    def _lookup_all(self, feats, dense):
        """Pooled output of every lookup group (the dense block rides with the first).  Groups share
        the ids, hence one backward sort (``PlanLink``) -- or one all-to-all routing when sharded."""
        if self._sharded is not None:
            return self._sharded(feats, dense)
        link = PlanLink() if self.training else None      # one sort per step, started early on a side stream
        outs = [self._groups[0](feats, dense, self.training, link)]
        for g in self._groups[1:]:
            outs.append(g(feats, None, self.training, link))
        return tuple(outs)

    # ---- optimizer / ids checks -------------------------------------------------------------
    def bind_optimizer(self, optimizer, kind: str | None = None):
        """Fuse the table update into backward, following ``optimizer``'s hyper-parameters (the
        optimizer itself keeps stepping the dense parameters; tables get no ``.grad``)."""
        if self._sharded is not None:
            self._sharded.bind_optimizer(optimizer, kind)
            return self
        for g in self._groups:
            g.bind_optimizer(optimizer, kind)
        return self

    def materialize_tables(self, device, seed: int | None = None, std: float = 1.0):
        """Allocate meta tables on ``device`` and fill them with the counter-based generator (the values a row-sharded copy
        of the model gets for the same ``seed``)."""
        from ..nn.embedding import EmbeddingTable
        seed = torch.initial_seed() if seed is None else seed
        groups = [g.tables for g in self._groups]
        for w, tables in enumerate(groups):
            for f, name in enumerate(self._names):
                t = tables[name]
                if t.weight.is_meta:
                    t.weight = nn.Parameter(torch.empty(t.num_embeddings, t.embedding_dim, dtype=torch.float32, device=device))
                t.counter_init_(EmbeddingTable.counter_seed(seed, f, w), std)
        return self

    def table_optimizer_state_dict(self):
        """Fused-update state that is NOT inside the torch optimizer (tables kept out of its parameters, row-wise Adagrad):
        {table module name: {"state0": tensor, "state1": tensor | None}}; plain tensors, ``weights_only`` safe.  State of
        tables that ARE parameters of the torch optimizer travels in ``optimizer.state_dict()`` instead."""
        from ..nn.embedding import EmbeddingTable
        out = {}
        for name, m in self.named_modules():
            if isinstance(m, EmbeddingTable) and m._state_owner is None and m._opt_state0 is not None:
                n = m.num_embeddings                                  # (the buffers may be sized for the table's reservation)
                out[name] = {"state0": m._opt_state0.detach()[:n].clone(),
                             "state1": None if m._opt_state1 is None else m._opt_state1.detach()[:n].clone()}
        return out

    def load_table_optimizer_state_dict(self, state):
        mods = dict(self.named_modules())
        for name, st in state.items():
            m = mods[name]
            m._opt_state0 = st["state0"].to(m.weight.device).clone()
            m._opt_state1 = None if st.get("state1") is None else st["state1"].to(m.weight.device).clone()

    def table_parameters(self):
        """Parameters updated by the fused sparse optimizer (every EmbeddingTable weight)."""
        from ..nn.embedding import EmbeddingTable
        params = [m.weight for m in self.modules() if isinstance(m, EmbeddingTable)]
        shards = getattr(self._sharded, "shards", None)
        if isinstance(shards, nn.ParameterList):          # peer-memory shards are bare parameters
            params += list(shards)
        return params

    def dense_parameters(self):
        """Everything else: hand these to the torch optimizer when the tables are fused, so that it
        does not allocate dense state for the tables."""
        skip = {id(p) for p in self.table_parameters()}
        return [p for p in self.parameters() if id(p) not in skip]

    def enable_flat_dense_grads(self):
        """Make the ``.grad`` of every dense parameter a view of ONE flat buffer (as DDP's buckets do): the multi-GPU
        all-reduce then needs no concatenation / copy-back.  Use ``zero_dense_grads()`` instead of
        ``optimizer.zero_grad(set_to_none=True)`` afterwards, which would drop the views."""
        params = [p for p in self.dense_parameters() if p.requires_grad]
        total = sum(p.numel() for p in params)
        flat = None
        adopt = getattr(self._sharded, "adopt_dense_grads", None)
        if adopt is not None:                 # hybrid placement: the buffer rides in the all-reduce of the replicated tables'
            flat = adopt(total)               # gradients at the end of backward; reduce_dense_grads() is then a no-op
            self._dense_grads_reduced_in_backward = flat is not None
        if flat is None:
            flat = torch.zeros(total, dtype=torch.float32, device=params[0].device)
        off = 0
        for p in params:
            p.grad = flat[off:off + p.numel()].view_as(p)
            p._ctr_direct_grad = True       # nn/tower.py: gradients made on the second stream are added into the view there
            off += p.numel()
        self._flat_dense_grad = flat
        return flat

    def zero_dense_grads(self):
        flat = getattr(self, "_flat_dense_grad", None)
        if flat is None:
            for p in self.dense_parameters():
                p.grad = None
                p._ctr_direct_grad = False
        else:
            flat.zero_()

    @staticmethod
    def dense_block(feats):
        parts = []
        if "dense_features" in feats:
            parts.append(feats["dense_features"])
        if "seq_dense_features" in feats:
            parts += list(feats["seq_dense_features"])       # dnn.py:64-65
        if not parts:
            return None
        return parts[0] if len(parts) == 1 else torch.cat([p.to(parts[0].device) for p in parts], dim=-1)

    def _first_linear(self, x: torch.Tensor, layer: nn.Linear) -> torch.Tensor:
        """Linear over the 4-float-padded lookup output: the weight is zero-padded instead of
        slicing (and copying) the activations."""
        pad = x.shape[1] - layer.in_features
        w = F.pad(layer.weight, (0, pad)) if pad else layer.weight
        return self._linear(x, w, layer.bias)

    @staticmethod
    def _linear(x, weight, bias):
        """Linear layer on the hand-written tcgen05 kernel: TF32 when torch is allowed to use TF32 for matmuls
        (``torch.backends.cuda.matmul.allow_tf32``), error-compensated 3xTF32 otherwise (the exact / parity mode)."""
        if x.is_cuda:
            return linear_tc(x, weight, bias)
        return F.linear(x, weight, bias)

    def _run_tower(self, x: torch.Tensor, stop_before_last: bool = False, blocked_ok: bool = False) -> torch.Tensor:
        """The tower of dnn.py:35-46 on the lookup output.  ``stop_before_last``: return the input of the final
        ``Linear(., 1)`` instead of the logits (the fused logit + loss head of ``training_step`` takes it from there).
        ``blocked_ok``: the caller states that ``x`` is the fused lookup's output and that NOTHING but this tower consumes it, so
        the first block may hand dL/dx back column-blocked (``nn.embedding.BlockedGrad``)."""
        layers = list(self.tower)
        if stop_before_last:
            layers = layers[:-1]
        if self.training and x.is_cuda and x.dtype == torch.float32:
            # training mode: every [Linear, BatchNorm1d, ReLU, Dropout] block is one fused autograd node
            prepared = self._take_prepared(x, layers)
            if prepared is not None:
                seed, prepared = prepared[0], prepared[1:]
            else:
                seed = self._step_seed(x.device)
            h, i, block = x, 0, 0
            while i + 3 < len(layers) and block_is_fusable(*layers[i:i + 4]):
                lin = layers[i]           # (first layer: the 4-float-padded lookup output; the block pads the weight itself)
                # sharded tables: let the first block write dL/dx straight into the peer-visible gradient matrix
                provider = getattr(self._sharded, "grad_buffer_provider", None) if block == 0 else None
                gx_provider = provider(0) if provider is not None else None
                if block == 0 and self._sharded is None:
                    offer = take_blocked_offer(x)          # (always taken, so that a stale offer never outlives its step)
                    gx_provider = offer if blocked_ok else None
                h = tower_block(h, lin, layers[i + 1], layers[i + 3], seed, block, gx_provider=gx_provider,
                                prepared=prepared if block == 0 else None)
                i += 4
                block += 1
            if i > 0:
                for layer in layers[i:]:
                    h = self._linear(h, layer.weight, layer.bias) if isinstance(layer, nn.Linear) else layer(h)
                return h
        if not layers:
            return x
        h = self._first_linear(x, layers[0])
        for layer in layers[1:]:
            h = self._linear(h, layer.weight, layer.bias) if isinstance(layer, nn.Linear) else layer(h)
        return h

    def _seed_counter(self, device):
        seed = getattr(self, "_drop_seed", None)
        if seed is None or seed.device != device:
            seed = torch.full((1,), torch.initial_seed() & 0x7fffffffffff, dtype=torch.int64, device=device)
            self._drop_seed = seed
        return seed

    def _prepare_tower(self):
        """Called by the models BEFORE they issue the lookup of a training forward: the first tower block's small preparatory
        kernels (dropout seed advance + snapshot, zero-padded weight, its transpose) go to the second stream now, so that they
        run next to the lookup instead of between it and the first GEMM (``nn.tower.prepare_first_block``).  Needs the padding
        the first block saw on an earlier step (``_run_tower`` records it); whatever is prepared here is taken -- or, if it does
        not fit, joined and dropped -- by the next ``_run_tower``."""
        self._prepared = None
        pad = getattr(self, "_first_block_pad", None)
        tower = getattr(self, "tower", None)
        if (pad is None or tower is None or not self.training or not torch.is_grad_enabled() or not _tower.prepare_early
                or not _tower.defer_weight_grads or len(tower) < 5 or not block_is_fusable(*list(tower)[:4])):
            return
        lin = tower[0]
        if not lin.weight.is_cuda or lin.weight.dtype != torch.float32:
            return
        precision = matmul_precision()
        with torch.cuda.device(lin.weight.device):
            made = prepare_first_block(lin, pad, self._seed_counter(lin.weight.device), want_wt=precision == "tf32")
        self._prepared = (made, pad, precision, lin.weight._version, lin)

    def _take_prepared(self, x, layers):
        """-> (seed snapshot, padded weight, transpose | None, event) when ``_prepare_tower`` ran for exactly this tower input,
        else None (anything prepared is joined first, so that no work is left dangling on the second stream)."""
        entry, self._prepared = getattr(self, "_prepared", None), None
        lin = layers[0] if layers else None
        fits = isinstance(lin, nn.Linear) and len(layers) >= 4 and block_is_fusable(*layers[:4])
        if fits:
            self._first_block_pad = x.shape[1] - lin.in_features
        if entry is None:
            return None
        made, pad, precision, version, prepared_lin = entry
        if (fits and prepared_lin is lin and pad == x.shape[1] - lin.in_features and precision == matmul_precision()
                and version == lin.weight._version and made[1].device == x.device):
            return made
        torch.cuda.current_stream(made[1].device).wait_event(made[3])
        # the counter was advanced by the preparation; the block falls back to a snapshot of its own (one more advance: harmless)
        return None

    def _step_seed(self, device):
        """Device-resident dropout seed, advanced once per training forward (inside a captured CUDA graph too)."""
        seed = self._seed_counter(device)
        seed += 1
        # a snapshot, not the counter itself: backward recomputes the dropout mask from the value ITS forward used, even
        # when another training forward ran in between (two losses summed, fwd-fwd-bwd orders)
        return seed.clone()

    def _grow_vocabularies(self, feats):
        if not self.training:
            return
        if self._sharded is not None:
            grow = getattr(self._sharded, "grow_vocabularies", None)
            if grow is not None and any(v is not None for v in getattr(self._sharded, "vocabs", [])):
                grow(feats)
            return
        for name in self._names:
            table = self.embeddings[name]
            if table.index_kind == "vocab":
                n = table.vocab.fit_and_grow(feats[name], table)
                twins = getattr(self, "linear_embeddings", None)
                if twins is not None and name in twins:   # DeepFM's first-order weights share the vocabulary
                    twins[name].grow_to(n)

    # ---- step protocol: dnn.py:72-82 ----------------------------------------------------------
    def hidden_and_extra(self, input_feats):
        """(input of the tower's last Linear [B, H], other logit terms [B, 1] or None): logits = last(h) + extra.
        Subclasses with extra logit terms (DeepFM) override."""
        raise NotImplementedError

    def training_step(self, batch, batch_idx):
        features, labels = batch
        last = self.tower[-1]
        if self.training and isinstance(last, nn.Linear) and len(self.tower) > 1:
            try:
                parts = self.hidden_and_extra(features)
            except NotImplementedError:
                parts = None
            if parts is not None:
                h, extra = parts[0], parts[1]
                second = parts[2] if len(parts) > 2 else None    # (xe, Linear): a linear term over raw features (DeepFM)
                labels = labels.to(h.device, non_blocking=True)
                if head_eligible(h, last, labels):       # last Linear + logit terms + mean BCE as one autograd node
                    if second is None:
                        return logit_bce(h, last, extra, labels)
                    if h.shape[1] >= 32:
                        return logit_bce(h, last, extra, labels, xe=second[0], linear_e=second[1])
                    term = second[1](second[0])
                    return logit_bce(h, last, term if extra is None else extra + term, labels)
                if second is not None:
                    term = second[1](second[0])
                    extra = term if extra is None else extra + term
                logits = self._linear(h, last.weight, last.bias)
                if extra is not None:
                    logits = logits + extra
                return F.binary_cross_entropy_with_logits(logits, labels)
        logits = self(features)
        return F.binary_cross_entropy_with_logits(logits, labels.to(logits.device, non_blocking=True))

    def validation_step(self, batch, batch_idx):
        features, labels = batch
        logits = self(features)
        return F.binary_cross_entropy_with_logits(logits, labels.to(logits.device, non_blocking=True))
