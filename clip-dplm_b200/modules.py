"""Drop-in counterparts of the reference's CLIP modules and loss functions.

Same class / function names, constructor arguments, forward signatures and output keys as the
reference, so that training scripts written against it keep working; only the tail of each forward
(normalise -> scaled similarity -> cross-entropy, and its backward) runs through the fused CUDA path.
The encoders / projection heads above that tail are ordinary torch modules -- they are not part of
the hot path and are kept structurally compatible (same parameter names, so reference state_dicts load).

    reference                                         here
    old/clip.py:38-73    RNAProteinCLIPModule          RNAProteinCLIPModule
    old/clip.py:75-110   DiffMapProteinCLIPModule      DiffMapProteinCLIPModule
    old/clip_opt.py:46   OptimizedCLIPModule           OptimizedCLIPModule   (+ optimized_clip_loss, :130-151)
    rna_clip_codes.ipynb:1925-1954  RNARBPCLIPModel    RNARBPCLIPModel
    tf_clip_codes (1).ipynb:13146-13165 (loss lines)   trimodal_contrastive_losses
    tong/utils/losses.py:4-19  contrastive_loss        contrastive_loss
    tong/utils/data.py:154-184 MemoryQueue             MemoryQueue

Differences a caller can observe, all deliberate:
  * output dicts gain "loss"; "logits_per_*" is a LazyLogits that only materialises the N x N matrix
    when touched (``.materialize()``, ``.argmax``, ``torch.*`` functions) -- large batches never allocate N^2;
  * with ``gather_distributed=True`` gradients flow through the gather (the reference's
    ``dist.all_gather`` + ``torch.cat`` silently cuts them, old/clip_opt.py:102-112).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from .engine import default_engine
from .functional import fused_clip_loss, fused_clip_loss_group

LOGIT_SCALE_INIT = math.log(1.0 / 0.07)   # 2.6592 (run1/configuration_hybrid_clip.py:100)


# ------------------------------------------------------------------------------------------------
# differentiable L2 normalise through the C-ABI (returned "*_embeds" stay part of the autograd graph)
# ------------------------------------------------------------------------------------------------
class _Normalize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        eng = default_engine()
        xc = x.detach().contiguous()
        rinv, xh = eng.normalize(xc, want_hat=xc.dtype)
        ctx.save_for_backward(xc, rinv)
        return xh

    @staticmethod
    def backward(ctx, g):
        xc, rinv = ctx.saved_tensors
        return default_engine().normalize_backward(xc, rinv, g.float().contiguous(), xc.dtype)


def fused_normalize(x: torch.Tensor) -> torch.Tensor:
    """F.normalize(x, dim=-1) for a [N,d] CUDA tensor (bf16 / fp32)."""
    return _Normalize.apply(x)


class LazyOutputs(dict):
    """The reference's output dict (old/clip.py:69-74) whose expensive entries are computed on first access; membership
    tests, ``keys()`` / ``items()`` / ``values()`` / ``get`` see every entry (listing the values materialises them)."""

    def __init__(self, eager=()):
        super().__init__(eager)
        self._makers = {}

    def lazy(self, key, maker):
        self._makers[key] = maker

    def __missing__(self, key):
        maker = self._makers.pop(key)        # KeyError for unknown keys, like a dict
        value = maker()
        self[key] = value
        return value

    def __contains__(self, key):
        return dict.__contains__(self, key) or key in self._makers

    def get(self, key, default=None):
        return self[key] if key in self else default

    def _all(self):
        for k in list(self._makers):
            self[k]
        return self

    def keys(self):
        return list(dict.keys(self)) + list(self._makers)

    def items(self):
        return dict.items(self._all())

    def values(self):
        return dict.values(self._all())

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return dict.__len__(self) + len(self._makers)


class LazyLogits:
    """Stands in for ``matmul(a, b.t()) * logit_scale`` (old/clip.py:67).  Nothing is computed until used; ``materialize``
    (any torch function applied to the stand-in) is a plain dense matmul for inspection and debugging, not a training path."""

    def __init__(self, a_hat, b_hat, scale):
        self.a_hat, self.b_hat, self.scale = a_hat, b_hat, scale
        self._m = None

    @property
    def shape(self):
        return torch.Size((self.a_hat.shape[0], self.b_hat.shape[0]))

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def materialize(self) -> torch.Tensor:
        if self._m is None:
            self._m = torch.matmul(self.a_hat.float(), self.b_hat.float().t()) * self.scale
        return self._m

    def argmax(self, dim=1):
        """``logits.argmax(dim=1)`` of run1/full.py:152 through the retrieval kernel (top-1 on the similarity sweep, no
        N x N matrix) where it is served: CUDA, d in {128, 256, 384, 512}; the positive scale does not move an argmax."""
        a, b = (self.a_hat, self.b_hat) if dim in (1, -1) else (self.b_hat, self.a_hat)
        if self._m is None and a.is_cuda and a.shape[1] % 128 == 0 and a.shape[1] <= 512 and dim in (0, 1, -1, -2):
            from .retrieval import topk_similarity
            return topk_similarity(a, b, 1)[1][:, 0]
        return self.materialize().argmax(dim=dim)

    def t(self):
        return LazyLogits(self.b_hat, self.a_hat, self.scale)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        conv = lambda o: o.materialize() if isinstance(o, LazyLogits) else o
        return func(*[conv(a) for a in args], **{k: conv(v) for k, v in (kwargs or {}).items()})


# ------------------------------------------------------------------------------------------------
# encoders / heads above the hot path (plain torch; parameter names match the reference)
# ------------------------------------------------------------------------------------------------
class CLIPEncoder(nn.Module):
    """``num_hidden_layers`` x (Linear + ReLU), then LayerNorm  (old/clip.py:8-17)."""

    def __init__(self, config):
        super().__init__()
        h = config.hidden_size
        self.layers = nn.ModuleList(nn.Linear(h, h) for _ in range(config.num_hidden_layers))
        self.layernorm = nn.LayerNorm(h, eps=config.layer_norm_eps)

    def forward(self, x):
        for lin in self.layers:
            x = torch.relu(lin(x))
        return self.layernorm(x)


def _mlp_head(dims, dropout):
    mods = []
    for k in range(len(dims) - 1):
        mods += [nn.Linear(dims[k], dims[k + 1]), nn.LayerNorm(dims[k + 1])]
        if k < len(dims) - 2:
            mods += [nn.GELU(), nn.Dropout(dropout)]
    return nn.Sequential(*mods)


class ProjectionHead(nn.Module):
    """Linear-LN-GELU-Dropout-Linear-LN  (old/clip.py:20-36).

    In bf16 (bf16 activations, or under autocast) on CUDA the last Linear -> LayerNorm pair runs as ONE tcgen05 kernel that
    also takes the row norms (heads.fused_linear_layernorm): the returned rows carry their 1/norm as ``_clipnce_rinv``,
    which the fused loss picks up instead of re-reading them.  fp32 models keep the reference's fp32 ops (``fuse_tail``
    switches the fusion off altogether)."""

    fuse_tail = True
    # measured on B200 (tools/bench_config2.py, 1280 -> 1024 -> 512 heads, whole step as one CUDA graph): at 4096 rows the
    # 32-CTA tail kernel loses to cuBLAS + two row kernels (0.61 vs 0.49 ms per step), at 65536 rows it wins (18.96 vs 19.56)
    fuse_tail_min_rows = 8192

    def __init__(self, input_dim, output_dim, hidden_dim=None, dropout=0.1):
        super().__init__()
        self.projection = _mlp_head([input_dim, hidden_dim or input_dim, output_dim], dropout)

    def forward(self, x):
        from .heads import fused_linear_layernorm, tail_is_served
        lin, ln = self.projection[-2], self.projection[-1]
        if (self.fuse_tail and x.is_cuda and x.dim() == 2 and x.shape[0] >= self.fuse_tail_min_rows
                and tail_is_served(lin.in_features, lin.out_features)
                and (x.dtype == torch.bfloat16 or torch.is_autocast_enabled())):
            h = x
            for m in list(self.projection)[:-2]:
                h = m(h)
            e, rinv = fused_linear_layernorm(h, lin, ln)
            e._clipnce_rinv = rinv
            return e
        return self.projection(x)


class OptimizedProjectionHead(nn.Module):
    """skip(x) + layer_scale * three-layer MLP, Xavier init  (old/clip_opt.py:9-44, notebook :1901-1909)."""

    def __init__(self, input_dim, output_dim, hidden_dim=None, dropout=0.1):
        super().__init__()
        hidden = hidden_dim or 2 * input_dim
        self.skip = nn.Linear(input_dim, output_dim)
        self.layer_scale = nn.Parameter(torch.full((1,), 1e-4))
        self.projection = _mlp_head([input_dim, hidden, hidden, output_dim], dropout)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                nn.init.zeros_(m.bias)

    def forward(self, x):
        return self.skip(x) + self.layer_scale * self.projection(x)


# ------------------------------------------------------------------------------------------------
# CLIP modules
# ------------------------------------------------------------------------------------------------
class _PairCLIP(nn.Module):
    """Shared body of the two-tower modules: encoders -> heads -> fused contrastive tail."""

    names = ("a", "b")      # output-key stems, set by subclasses
    symmetric = True
    clamp_max: Optional[float] = None

    # with a process group: the module is assumed to be wrapped in DistributedDataParallel, as the reference does
    # (old/clip_opt.py:154, run1/full.py:172) -- row gradients follow DDP's averaging convention (fused_clip_loss `ddp`)
    ddp_gradients = True

    def _tail(self, emb_a, emb_b, extra_cols=None, group=None):
        loss = fused_clip_loss(emb_a, emb_b, self.logit_scale, symmetric=self.symmetric, clamp_max=self.clamp_max,
                               extra_cols=extra_cols, group=group, ddp=self.ddp_gradients and group is not None,
                               rinv_a=getattr(emb_a, "_clipnce_rinv", None), rinv_b=getattr(emb_b, "_clipnce_rinv", None))
        na, nb = self.names

        def scale():
            s = self.logit_scale.detach().exp()
            return s.clamp(max=self.clamp_max) if self.clamp_max is not None else s

        # the normalised embeddings and the logits stand-in are only formed if somebody asks for them: a training loop
        # that reads outputs["loss"] pays for no pass over [N, d] beyond the fused loss itself
        out = LazyOutputs({"loss": loss})
        out.lazy(f"{na}_embeds", lambda: fused_normalize(emb_a))
        out.lazy(f"{nb}_embeds", lambda: fused_normalize(emb_b))
        out.lazy(f"logits_per_{na}_{nb}", lambda: LazyLogits(out[f"{na}_embeds"].detach(), out[f"{nb}_embeds"].detach(), scale()))
        return out


class RNAProteinCLIPModule(_PairCLIP):
    names = ("rna", "protein")

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.rna_model = CLIPEncoder(config.rna_config)
        self.protein_model = CLIPEncoder(config.protein_config)
        p = config.projection_dim
        self.rna_projection = ProjectionHead(config.rna_config.hidden_size, p, hidden_dim=2 * p)
        self.protein_projection = ProjectionHead(config.protein_config.hidden_size, p, hidden_dim=2 * p)
        self.logit_scale = nn.Parameter(torch.ones([]) * config.logit_scale_init_value)

    def forward(self, rna_values, protein_values):
        return self._tail(self.rna_projection(self.rna_model(rna_values)),
                          self.protein_projection(self.protein_model(protein_values)))


class DiffMapProteinCLIPModule(_PairCLIP):
    names = ("diffmap", "protein")

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.diffmap_model = CLIPEncoder(config.diffmap_config)
        self.protein_model = CLIPEncoder(config.protein_config)
        p = config.projection_dim
        self.diffmap_projection = ProjectionHead(config.diffmap_config.hidden_size, p, hidden_dim=2 * p)
        self.protein_projection = ProjectionHead(config.protein_config.hidden_size, p, hidden_dim=2 * p)
        self.logit_scale = nn.Parameter(torch.ones([]) * config.logit_scale_init_value)

    def forward(self, diffmap_values, protein_values):
        return self._tail(self.diffmap_projection(self.diffmap_model(diffmap_values)),
                          self.protein_projection(self.protein_model(protein_values)))


class OptimizedCLIPModule(_PairCLIP):
    """old/clip_opt.py:46-128: wider skip heads, FIFO cache of normalised protein embeddings used as extra
    negative columns (:52-56, :76-81, :118-121), ``exp().clamp(max=100)`` (:100), global negatives (:102-112)."""

    names = ("diffmap", "protein")
    clamp_max = 100.0

    def __init__(self, config):
        super().__init__()
        self.config = config
        p = config.projection_dim
        self.register_buffer("protein_embedding_cache", torch.zeros(config.cache_size, p), persistent=False)
        self.cache_ptr = 0
        self.diffmap_model = CLIPEncoder(config.diffmap_config)
        self.protein_model = CLIPEncoder(config.protein_config)
        self.diffmap_projection = OptimizedProjectionHead(config.diffmap_config.hidden_size, p, hidden_dim=4 * p)
        self.protein_projection = OptimizedProjectionHead(config.protein_config.hidden_size, p, hidden_dim=4 * p)
        self.logit_scale = nn.Parameter(torch.ones([]) * LOGIT_SCALE_INIT)

    @torch.no_grad()
    def update_cache(self, protein_embeds):
        """FIFO write that restarts at 0 when the batch does not fit (old/clip_opt.py:76-81)."""
        n = protein_embeds.size(0)
        if self.cache_ptr + n > self.config.cache_size:
            self.cache_ptr = 0
        self.protein_embedding_cache[self.cache_ptr:self.cache_ptr + n] = protein_embeds.detach().to(
            self.protein_embedding_cache.dtype)
        self.cache_ptr = (self.cache_ptr + n) % self.config.cache_size

    def forward(self, diffmap_values, protein_values, gather_distributed=True):
        ea = self.diffmap_projection(self.diffmap_model(diffmap_values))
        eb = self.protein_projection(self.protein_model(protein_values))
        self.update_cache(fused_normalize(eb.detach()))       # the reference updates before the similarity (:97)
        cache = self.protein_embedding_cache[:self.cache_ptr]
        group = dist.group.WORLD if (gather_distributed and dist.is_available() and dist.is_initialized()) else None
        out = self._tail(ea, eb, extra_cols=cache if cache.shape[0] else None, group=group)
        out.lazy("logits_per_diffmap_cache", lambda: LazyLogits(out["diffmap_embeds"].detach(), cache,
                                                                 self.logit_scale.detach().exp().clamp(max=100.0)))
        return out


def optimized_clip_loss(outputs, temperature=0.07):
    """old/clip_opt.py:130-151 -- (CE([S | S_cache]) + CE(S^T)) / 2 (its label-smoothing tensors are dead code and
    ``temperature`` is unused there too).  The fused modules already computed it."""
    if "loss" in outputs:
        return outputs["loss"]
    raise RuntimeError("optimized_clip_loss expects the outputs of a clip_dplm_b200 module (they carry 'loss'); "
                       "materialised logits are never re-read on the product path")


# ------------------------------------------------------------------------------------------------
# RNA <-> RBP model of current/rna_clip_codes.ipynb (cells 24 + 28)
# ------------------------------------------------------------------------------------------------
def create_padding_mask(emb):
    """True where a position holds data; padding rows are NaN-filled (rna_clip_codes.ipynb:1841-1845)."""
    return ~torch.isnan(emb).any(dim=-1)


class RNARBPCLIPEncoder(nn.Module):
    def __init__(self, embed_dim, num_layers=3):
        super().__init__()
        self.layers = nn.ModuleList(nn.TransformerEncoderLayer(d_model=embed_dim, nhead=8, dim_feedforward=4 * embed_dim,
                                                               dropout=0.1) for _ in range(num_layers))
        self.layernorm = nn.LayerNorm(embed_dim)

    def forward(self, x, src_key_padding_mask=None):
        for layer in self.layers:
            x = layer(x, src_key_padding_mask=src_key_padding_mask)
        return self.layernorm(x)


class RNARBPCLIPProjectionHead(OptimizedProjectionHead):
    def __init__(self, input_dim, output_dim):
        nn.Module.__init__(self)
        self.skip = nn.Linear(input_dim, output_dim)
        self.layer_scale = nn.Parameter(torch.ones(1) * 1e-4)
        self.projection = _mlp_head([input_dim, 2 * input_dim, 2 * input_dim, output_dim], 0.1)


class RNARBPCLIPModel(nn.Module):
    """forward(rna_emb, rbp_emb) -> (rna_embed, rbp_embed, loss)   (rna_clip_codes.ipynb:1935-1954)."""

    def __init__(self, rna_dim=120, rbp_dim=1280, projection_dim=512):
        super().__init__()
        self.rna_encoder = RNARBPCLIPEncoder(rna_dim)
        self.rbp_encoder = RNARBPCLIPEncoder(rbp_dim)
        self.rna_projection = RNARBPCLIPProjectionHead(rna_dim, projection_dim)
        self.rbp_projection = RNARBPCLIPProjectionHead(rbp_dim, projection_dim)
        self.logit_scale = nn.Parameter(torch.ones([]) * LOGIT_SCALE_INIT)

    def forward(self, rna_emb, rbp_emb):
        rna_mask = create_padding_mask(rna_emb).transpose(0, 1)
        rbp_mask = create_padding_mask(rbp_emb).transpose(0, 1)
        rna_emb = torch.nan_to_num(rna_emb, 0.0)
        rbp_emb = torch.nan_to_num(rbp_emb, 0.0)
        rna_enc = self.rna_encoder(rna_emb, src_key_padding_mask=~rna_mask)
        rbp_enc = self.rbp_encoder(rbp_emb, src_key_padding_mask=~rbp_mask)
        pa, pb = self.rna_projection(rna_enc[:, 0]), self.rbp_projection(rbp_enc[:, 0])
        loss = fused_clip_loss(pa, pb, self.logit_scale)          # replaces notebook lines 1948-1953
        return fused_normalize(pa), fused_normalize(pb), loss


# ------------------------------------------------------------------------------------------------
# tri-modal losses and the tong queue variant
# ------------------------------------------------------------------------------------------------
def trimodal_contrastive_losses(cell_embed, pert_embed, protein_embed, logit_scale):
    """Three pairwise symmetric InfoNCE losses sharing one logit_scale, summed
    (tf_clip_codes (1).ipynb:13146-13165).  Inputs are the three projection outputs (un-normalised)."""
    # every embedding is an operand of two pairs: its row norms are taken ONCE (or arrive with it from the fused head
    # tail) and handed to both pair steps; the normalised copies of the reference's output dict are formed on demand
    embs = (cell_embed, pert_embed, protein_embed)
    # one launch per kernel for the three pairs where the kernels serve the shapes (functional.fused_clip_loss_group):
    # one normalise pass over the stacked embeddings, one forward sweep, one backward sweep over all six sides
    holder = {}
    losses = fused_clip_loss_group(embs, (0, 0, 1), (1, 2, 2), logit_scale, holder=holder)
    if losses is not None:
        out = LazyOutputs({"loss": losses[3], "cell_pert_loss": losses[0], "cell_protein_loss": losses[1],
                           "pert_protein_loss": losses[2]})
        out.grad_info = holder      # after backward: holder["embed_grad_sumsq"] = |d cell|^2, |d pert|^2, |d protein|^2
    else:
        if all(e.is_cuda for e in embs):
            eng = default_engine()
            rinv = [getattr(e, "_clipnce_rinv", None) for e in embs]
            rinv = [r if r is not None else eng.normalize(e.detach().contiguous() if e.dtype != torch.float16
                                                          else e.detach().to(torch.bfloat16).contiguous())[0]
                    for r, e in zip(rinv, embs)]
        else:
            rinv = [None, None, None]
        cp = fused_clip_loss(cell_embed, pert_embed, logit_scale, rinv_a=rinv[0], rinv_b=rinv[1])
        cq = fused_clip_loss(cell_embed, protein_embed, logit_scale, rinv_a=rinv[0], rinv_b=rinv[2])
        pq = fused_clip_loss(pert_embed, protein_embed, logit_scale, rinv_a=rinv[1], rinv_b=rinv[2])
        out = LazyOutputs({"loss": cp + cq + pq, "cell_pert_loss": cp, "cell_protein_loss": cq, "pert_protein_loss": pq})
    out.lazy("cell_embed", lambda: fused_normalize(cell_embed))
    out.lazy("pert_embed", lambda: fused_normalize(pert_embed))
    out.lazy("protein_embed", lambda: fused_normalize(protein_embed))
    return out


def contrastive_loss(x, y, temperature=0.1, queue=None):
    """tong/utils/losses.py:4-19 -- one-directional InfoNCE, logits divided by a fixed temperature, optional memory
    queue appended to the columns as stored (no gradient)."""
    return fused_clip_loss(x, y, 1.0 / temperature, symmetric=False, scale_is_log=False, extra_cols=queue,
                           extra_normalized=False if queue is not None else True)


class MemoryQueue:
    """FIFO of detached embeddings with split write on wrap-around (tong/utils/data.py:154-184)."""

    def __init__(self, size, dim, device=None, dtype=torch.float32):
        self.size, self.dim, self.ptr = size, dim, 0
        self.queue = torch.zeros(size, dim, device=device, dtype=dtype)

    @torch.no_grad()
    def enqueue_dequeue(self, embeddings):
        e = embeddings.detach().to(self.queue.dtype)
        n = e.shape[0]
        if self.ptr + n > self.size:
            first = self.size - self.ptr
            self.queue[self.ptr:] = e[:first]
            self.queue[:n - first] = e[first:]
            self.ptr = n - first
        else:
            self.queue[self.ptr:self.ptr + n] = e
            self.ptr = (self.ptr + n) % self.size
        return self.queue
