"""Drop-in modules (clip_dplm_b200.modules) against the reference math, on a B200."""
import math
import types

import pytest
import torch

from oracle import ref_step as O

pytestmark = pytest.mark.gpu


def cfg(ha, hb, proj, cache=64):
    sub = lambda h: types.SimpleNamespace(hidden_size=h, num_hidden_layers=2, layer_norm_eps=1e-5)
    return types.SimpleNamespace(rna_config=sub(ha), protein_config=sub(hb), diffmap_config=sub(ha), projection_dim=proj,
                                 logit_scale_init_value=2.6592, cache_size=cache)


def rel(x, ref):
    x, ref = x.double().cpu(), ref.double().cpu()
    return float((x - ref).norm() / ref.norm().clamp_min(1e-300))


def test_rna_protein_module_matches_reference_tail():
    from clip_dplm_b200 import modules as M
    torch.manual_seed(0)
    m = M.RNAProteinCLIPModule(cfg(48, 64, 128)).cuda().eval()
    x, y = torch.randn(96, 48, device="cuda"), torch.randn(96, 64, device="cuda")
    out = m(x, y)
    ea, eb = m.rna_projection(m.rna_model(x)), m.protein_projection(m.protein_model(y))
    ref_logits, ah, bh = O.ref_logits(ea.detach().cpu().double(), eb.detach().cpu().double(), m.logit_scale.detach().cpu().double())
    ref_loss = O.ref_loss(ea.detach().cpu().double(), eb.detach().cpu().double(), m.logit_scale.detach().cpu().double())
    assert set(out) == {"logits_per_rna_protein", "rna_embeds", "protein_embeds", "loss"}
    assert rel(out["rna_embeds"], ah) < 1e-5 and rel(out["protein_embeds"], bh) < 1e-5
    assert abs(float(out["loss"]) - float(ref_loss)) <= 1e-5 * abs(float(ref_loss))
    lg = out["logits_per_rna_protein"]
    assert tuple(lg.shape) == (96, 96)
    assert rel(lg.materialize(), ref_logits) < 1e-5
    assert torch.equal(lg.argmax(dim=1).cpu(), ref_logits.argmax(dim=1))
    # the loss trains every parameter above the tail, and the returned embeds stay differentiable
    (out["loss"] + out["rna_embeds"].sum() * 1e-3).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())


def test_gradients_through_module_match_autograd_of_reference():
    from clip_dplm_b200 import modules as M
    torch.manual_seed(1)
    m = M.DiffMapProteinCLIPModule(cfg(32, 40, 64)).cuda().eval()
    x, y = torch.randn(80, 32, device="cuda"), torch.randn(80, 40, device="cuda")
    m(x, y)["loss"].backward()
    g_fused = {k: p.grad.clone() for k, p in m.named_parameters()}
    m.zero_grad()
    ea, eb = m.diffmap_projection(m.diffmap_model(x)), m.protein_projection(m.protein_model(y))
    O.ref_loss(ea, eb, m.logit_scale).backward()          # the reference's op sequence on the GPU, fp32 autograd
    for k, p in m.named_parameters():
        assert rel(g_fused[k], p.grad) < 2e-4, k


def test_optimized_module_cache_and_clamp():
    from clip_dplm_b200 import modules as M
    torch.manual_seed(2)
    c = cfg(32, 32, 64, cache=96)
    m = M.OptimizedCLIPModule(c).cuda().eval()
    x, y = torch.randn(48, 32, device="cuda"), torch.randn(48, 32, device="cuda")
    out1 = m(x, y, gather_distributed=False)
    assert m.cache_ptr == 48 and out1["logits_per_diffmap_cache"].shape == (48, 48)
    x2, y2 = torch.randn(48, 32, device="cuda"), torch.randn(48, 32, device="cuda")
    out2 = m(x2, y2, gather_distributed=False)        # cache now holds batch 1 and batch 2 (wraps to 0 after 96)
    assert m.cache_ptr == 0 and out2["logits_per_diffmap_cache"].shape == (48, 0) or m.cache_ptr in (0, 96)
    ea, eb = m.diffmap_projection(m.diffmap_model(x)), m.protein_projection(m.protein_model(y))
    cache = torch.nn.functional.normalize(eb.detach(), dim=-1)
    ref = O.ref_loss(ea.detach().cpu().double(), eb.detach().cpu().double(), m.logit_scale.detach().cpu().double(),
                     clamp_max=100, extra_cols=cache.cpu().double())
    assert abs(float(out1["loss"]) - float(ref)) <= 1e-4 * abs(float(ref))
    assert float(M.optimized_clip_loss(out1)) == float(out1["loss"])


def test_notebook_model_and_tong_loss():
    from clip_dplm_b200 import modules as M
    torch.manual_seed(3)
    model = M.RNARBPCLIPModel(rna_dim=16, rbp_dim=24, projection_dim=32).cuda().eval()
    rna, rbp = torch.randn(12, 5, 16, device="cuda"), torch.randn(12, 7, 24, device="cuda")
    ra, rb, loss = model(rna, rbp)
    assert ra.shape == (12, 32) and rb.shape == (12, 32)
    assert abs(float(ra.norm(dim=1).mean()) - 1) < 1e-4
    with torch.no_grad():
        re_ = model.rna_encoder(rna, src_key_padding_mask=~M.create_padding_mask(rna).transpose(0, 1))
        be_ = model.rbp_encoder(rbp, src_key_padding_mask=~M.create_padding_mask(rbp).transpose(0, 1))
        pa, pb = model.rna_projection(re_[:, 0]), model.rbp_projection(be_[:, 0])
    ref_l = O.ref_loss(pa.cpu().double(), pb.cpu().double(), model.logit_scale.detach().cpu().double())
    assert abs(float(loss) - float(ref_l)) <= 1e-5 * abs(float(ref_l))
    x, y, q = torch.randn(33, 40, device="cuda"), torch.randn(33, 40, device="cuda"), torch.randn(17, 40, device="cuda")
    l = M.contrastive_loss(x, y, temperature=0.1, queue=q)
    ref = O.ref_loss(x.cpu().double(), y.cpu().double(), torch.tensor(10.0, dtype=torch.float64), symmetric=False,
                     scale_is_log=False, extra_cols=q.cpu().double())
    assert abs(float(l) - float(ref)) <= 1e-5 * abs(float(ref))
    tri = M.trimodal_contrastive_losses(x, y, torch.randn(33, 40, device="cuda"), torch.tensor(2.0, device="cuda"))
    assert abs(float(tri["loss"]) - float(tri["cell_pert_loss"] + tri["cell_protein_loss"] + tri["pert_protein_loss"])) < 1e-5


# ------------------------------------------------------------------------------------------------ projection-head tail
@pytest.mark.gpu
@pytest.mark.parametrize("n,k,p", [(1000, 1024, 512), (4096, 1280, 512), (300, 256, 128), (513, 640, 384)])
def test_fused_head_tail_matches_linear_layernorm_normalize(n, k, p):
    """heads.fused_linear_layernorm (one tcgen05 kernel: Linear -> LayerNorm -> row norm) against the reference's own
    ops on the same bf16-rounded operands (old/clip.py:26-33 last two layers + F.normalize :63-64), forward and backward."""
    import torch.nn.functional as F
    from clip_dplm_b200.heads import fused_linear_layernorm
    torch.manual_seed(7)
    lin = torch.nn.Linear(k, p).cuda()
    ln = torch.nn.LayerNorm(p).cuda()
    with torch.no_grad():
        ln.weight.uniform_(0.5, 1.5)
        ln.bias.uniform_(-0.3, 0.3)
        lin.weight.copy_(lin.weight.bfloat16().float())
    h = torch.randn(n, k, device="cuda").bfloat16().requires_grad_(True)
    e, rinv = fused_linear_layernorm(h, lin, ln)
    href = h.detach().double().requires_grad_(True)
    lin64, ln64 = torch.nn.Linear(k, p).cuda().double(), torch.nn.LayerNorm(p).cuda().double()
    lin64.load_state_dict({kk: v.double() for kk, v in lin.state_dict().items()})
    ln64.load_state_dict({kk: v.double() for kk, v in ln.state_dict().items()})
    eref = ln64(lin64(href))
    assert rel(e, eref) <= 5e-3                                           # bf16 output rounding
    assert torch.allclose(rinv.double(), 1.0 / e.double().norm(dim=1).clamp_min(1e-12), rtol=1e-5)
    assert rel(e.double() * rinv.double()[:, None], F.normalize(eref, dim=-1)) <= 5e-3
    g = torch.randn(n, p, device="cuda")
    e.backward(g.bfloat16())
    eref.backward(g.bfloat16().double())
    assert rel(h.grad, href.grad) <= 2e-2
    assert rel(lin.weight.grad, lin64.weight.grad) <= 2e-2 and rel(lin.bias.grad, lin64.bias.grad) <= 2e-2
    assert rel(ln.weight.grad, ln64.weight.grad) <= 2e-2 and rel(ln.bias.grad, ln64.bias.grad) <= 2e-2


@pytest.mark.gpu
def test_module_with_fused_tail_matches_unfused_module():
    """RNAProteinCLIPModule in bf16: heads with the fused tail (rows + 1/norm handed straight to the loss) against the same
    module with fuse_tail off -- loss and every parameter gradient; and the lazily formed output entries."""
    from clip_dplm_b200 import modules as M
    c = cfg(96, 160, 256)
    torch.manual_seed(3)
    m1 = M.RNAProteinCLIPModule(c).cuda().bfloat16()
    m2 = M.RNAProteinCLIPModule(c).cuda().bfloat16()
    m2.load_state_dict(m1.state_dict())
    for mod in m1.modules():
        if isinstance(mod, M.ProjectionHead):
            mod.fuse_tail_min_rows = 1          # the default only fuses from 8192 rows on
    for mod in m2.modules():
        if isinstance(mod, M.ProjectionHead):
            mod.fuse_tail = False
    for mod in list(m1.modules()) + list(m2.modules()):
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    xa, xb = torch.randn(640, 96, device="cuda").bfloat16(), torch.randn(640, 160, device="cuda").bfloat16()
    o1, o2 = m1(xa, xb), m2(xa, xb)
    assert set(o1.keys()) == {"loss", "rna_embeds", "protein_embeds", "logits_per_rna_protein"} and "rna_embeds" in o1
    o1["loss"].backward()
    o2["loss"].backward()
    assert abs(float(o1["loss"]) - float(o2["loss"])) <= 2e-2 * abs(float(o2["loss"]))
    assert rel(o1["rna_embeds"], o2["rna_embeds"]) <= 2e-2
    for (n1, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        if p2.grad is not None and float(p2.grad.float().norm()) > 0:
            assert rel(p1.grad, p2.grad) <= 6e-2, n1        # two bf16 models: rounding of every intermediate differs


# ------------------------------------------------------------------------------------------------ tri-modal grouped launch
def _tri_inputs(n, d, seed, mixes=(1.0, 0.5, 0.3)):
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(n, d, generator=g)
    mk = lambda mix: (mix * base + (1.0 - mix) * torch.randn(n, d, generator=g)).bfloat16().float()
    return tuple(mk(m) for m in mixes)


@pytest.mark.parametrize("n,d,t,dtype", [(256, 128, 2.6592, torch.bfloat16), (1000, 512, 2.6592, torch.bfloat16),
                                         (4096, 512, 2.6592, torch.bfloat16), (777, 768, 2.6592, torch.bfloat16),
                                         (640, 256, math.log(100.0), torch.bfloat16), (1536, 384, 2.0, torch.float32)])
def test_trimodal_grouped_launch_matches_three_reference_pairs(n, d, t, dtype):
    """modules.trimodal_contrastive_losses through the grouped launch (clipnce_group_*: three pairs in one forward sweep,
    their six backward sides in one sweep, one finishing pass) against three independent evaluations of the reference's
    loss lines (tf_clip_codes (1).ipynb:13146-13165) in float64 -- losses, all three embedding gradients (each the sum of
    two pairs' contributions), d logit_scale, with DIFFERENT upstream weights on the three losses and on their sum."""
    from clip_dplm_b200 import modules as M
    from clip_dplm_b200 import functional as Fn
    # at s = 100 weakly correlated members keep the losses (and gradients) away from zero
    c, p, q = _tri_inputs(n, d, 11 + n, mixes=(1.0, 0.5, 0.3) if t < 4.0 else (1.0, 0.1, 0.05))
    w = [0.7, 1.3, -0.4, 0.9]     # weights of cell_pert, cell_protein, pert_protein, total

    # reference: float64 autograd on the same bf16-rounded values
    rc, rp, rq = (x.double().requires_grad_(True) for x in (c, p, q))
    rt = torch.tensor(t, dtype=torch.float64, requires_grad=True)
    l_cp, l_cq, l_pq = O.ref_loss(rc, rp, rt), O.ref_loss(rc, rq, rt), O.ref_loss(rp, rq, rt)
    (w[0] * l_cp + w[1] * l_cq + w[2] * l_pq + w[3] * (l_cp + l_cq + l_pq)).backward()

    embs = [x.cuda().to(dtype).requires_grad_(True) for x in (c, p, q)]
    ls = torch.tensor(t, device="cuda", requires_grad=True)
    calls = []
    orig = Fn._GroupedClipLoss.apply
    Fn._GroupedClipLoss.apply = staticmethod(lambda *a: (calls.append(1), orig(*a))[1])
    try:
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=dtype == torch.float32):
            out = M.trimodal_contrastive_losses(embs[0], embs[1], embs[2], ls)
    finally:
        Fn._GroupedClipLoss.apply = orig
    assert calls, "the grouped launch did not serve this shape"
    for key, ref in (("cell_pert_loss", l_cp), ("cell_protein_loss", l_cq), ("pert_protein_loss", l_pq),
                     ("loss", l_cp + l_cq + l_pq)):
        assert abs(float(out[key]) - float(ref)) <= 1e-3 * abs(float(ref)) + 1e-6, key
    (w[0] * out["cell_pert_loss"] + w[1] * out["cell_protein_loss"] + w[2] * out["pert_protein_loss"]
     + w[3] * out["loss"]).backward()
    torch.cuda.synchronize()
    for e, r, name in zip(embs, (rc, rp, rq), ("cell", "pert", "protein")):
        assert e.grad.dtype == dtype and rel(e.grad, r.grad) <= 2e-2, (name, rel(e.grad, r.grad))
    assert abs(float(ls.grad) - float(rt.grad)) <= 2e-2 * abs(float(rt.grad)) + 1e-5
    sq = out.grad_info["embed_grad_sumsq"].cpu()
    for i, e in enumerate(embs):
        assert abs(float(sq[i]) - float(e.grad.double().pow(2).sum())) <= 2e-2 * float(sq[i])
    assert rel(out["pert_embed"], torch.nn.functional.normalize(p.double(), dim=-1)) <= 5e-3


def test_trimodal_grouped_launch_is_deterministic_and_graph_capturable():
    from clip_dplm_b200 import modules as M
    c, p, q = (x.cuda().bfloat16() for x in _tri_inputs(2048, 512, 5))
    ls = torch.tensor(2.6592, device="cuda", requires_grad=True)

    def step():
        embs = [x.clone().requires_grad_(True) for x in (c, p, q)]
        out = M.trimodal_contrastive_losses(*embs, ls)
        out["loss"].backward()
        return out["loss"].detach().clone(), [e.grad for e in embs]

    l0, g0 = step()
    l1, g1 = step()
    assert torch.equal(l0, l1) and all(torch.equal(a, b) for a, b in zip(g0, g1))
    # capture: static inputs, one replay = three pairs forward + backward
    se = [x.clone().requires_grad_(True) for x in (c, p, q)]
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(2):
            o = M.trimodal_contrastive_losses(*se, ls)
            o["loss"].backward()
    torch.cuda.current_stream().wait_stream(side)
    for e in se:
        e.grad = None
    ls.grad = None
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        o = M.trimodal_contrastive_losses(*se, ls)
        o["loss"].backward()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(o["loss"].detach(), l0) and all(torch.equal(e.grad, g) for e, g in zip(se, g0))
