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
