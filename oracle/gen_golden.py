"""TEST INFRASTRUCTURE — pins `oracle/restated.py` against the UNMODIFIED reference and writes
`tests/golden/*.npz`.

Run in the build container only (needs /root/reference):   python -m oracle.gen_golden

What is executed is the reference's own code: `models/dit.py` is imported as-is (two dependency shims,
see ref_loader.py) and the `Diffusion` methods `q_xt`, `_sample_t`, `_subs_parameterization`, `forward`,
`compute_loss` (model.py) plus `_sample_categorical`, `_ddpm_forward`, `_ddpm_update`,
`_ddpm_caching_update`, `adap_sche`, `_maskgit_update` (model_utils.py / model_eval.py) are pulled out of
the source files with `ast` and exec'd against a minimal fake `self`.  Each block below (1) runs the
reference, (2) asserts the restatement matches (bit-exact for integer outputs, tight tolerance for fp32),
(3) stores inputs + reference outputs as fixtures.
"""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader as RL  # noqa: E402
from oracle import restated as R  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _np(t):
    if isinstance(t, torch.Tensor):
        if t.dtype == torch.bfloat16:
            return t.float().numpy()
        return t.detach().cpu().numpy()
    return np.asarray(t)


def make_fake_self(cfgd, ref_dit=None, *, training=True, mask_entire_modality=None, img_loss_weight=0.6,
                   softmin_snr=None, eval_cfg=None):
    noise_mod = RL.load_reference_noise()
    M = RL.load_reference_diffusion_methods()
    s = SimpleNamespace()
    s.config = RL.to_attrdict(dict(
        backbone="dit", mode="train", parameterization="subs",
        trainer=dict(mask_entire_modality=mask_entire_modality, multimodal_batches=True, interleaved=False,
                     joint_ar_nar_prob=None, add_label=False, first_token_dropout=None,
                     joint_ar_nar_timestep_warmup_steps=None, force_timestep=None, image_mode="discrete",
                     ar_llm_loss=False, allow_null_sigma=True, force_null_sigma=True,
                     disable_forward_autocast_during_eval=False, compile=False, force_bf16_eval=False,
                     ar_shift=False, low_precision_loss=False, log_seperate_modal_losses=True,
                     text_loss_weight=1.0, img_loss_weight=img_loss_weight, softmin_snr=softmin_snr,
                     interleaved_training_flex_attention=False),
        model=dict(force_argmax_valid_indices=True, use_attention_mask=False, length=cfgd["txt"] + cfgd["img"],
                   flex_attention_img_masking_prob=None, flex_attention_txt_masking_prob=None,
                   txt_length=cfgd["txt"], img_length=cfgd["img"]),
        eval=dict(ar_inpainting_force_val=None, cfg=eval_cfg),
        noise=dict(type="loglinear"),
    ))
    s.backbone = ref_dit if ref_dit is not None else SimpleNamespace(training=training)
    s.training = training
    s.mask_index = cfgd["mask_index"]
    s.text_vocab_size = cfgd["text_vocab_size"]
    s.vocab_size = cfgd["vocab_size"]
    s.parameterization = "subs"
    s.antithetic_sampling = True
    s.importance_sampling = False
    s.change_of_variables = False
    s.sampling_eps = 1e-3
    s.allow_slicing = False
    s.neg_infinity = -1_000_000.0
    s.time_conditioning = False
    s.T = 0
    s.dtype = torch.float32
    s.device = torch.device("cpu")
    s.is_compiled = True          # skips utils.print_nans (model.py:943)
    s.current_run_fwd_bwd_pass = 1
    s.noise = noise_mod.LogLinearNoise()
    s.global_step = 0
    s.visualize_samples = lambda *a, **k: None
    s._maybe_sub_sample = lambda x0, am: (x0, None, am)
    for name in ["q_xt", "_sample_t", "_subs_parameterization", "_process_sigma", "_ddpm_forward", "_ddpm_update",
                 "_ddpm_caching_update", "get_cfg_weight", "_maskgit_update", "_sample_prior"]:
        fn = getattr(M, name)
        setattr(s, name, (lambda f: (lambda *a, **k: f(s, *a, **k)))(fn))
    return s, M


def gen_dit():
    """Backbone: reference DIT (fp32 and CPU-bf16-autocast) vs restatement; stores params+io."""
    cfgd = dict(D=128, H=2, L=2, txt=64, img=64, text_vocab_size=97, vocab_size=160, mask_index=96)
    ref_cfg = RL.make_ref_config(cfgd["D"], cfgd["H"], cfgd["L"], cfgd["txt"], cfgd["img"])
    torch.manual_seed(0)
    dit = RL.build_reference_dit(ref_cfg, cfgd["vocab_size"], cfgd["text_vocab_size"], cfgd["mask_index"], dtype=torch.float32)
    dit.eval()
    # perturb norm weights / LN bias away from the (1, 0) init so that they are actually exercised
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for n, p in dit.named_parameters():
            if "norm" in n and n.endswith("weight"):
                p.add_((torch.rand(p.shape, generator=g) - 0.5) * 0.4)
            if "norm" in n and n.endswith("bias"):
                p.add_((torch.rand(p.shape, generator=g) - 0.5) * 0.2)
            if n == "output_layer.linear.bias":
                p.add_((torch.rand(p.shape, generator=g) - 0.5) * 0.2)
    P = {k: v.detach().clone() for k, v in dit.state_dict().items()}
    ids, modality = R.synthetic_batch(2, cfgd["txt"], cfgd["img"], cfgd["text_vocab_size"], cfgd["vocab_size"], seed=42)
    ids_clean = ids.clone()
    ids[0, 3] = cfgd["mask_index"]
    ids[1, 70:90] = cfgd["mask_index"]
    with torch.no_grad():
        ref_logits = dit(ids, None, modality=modality)
    ocfg = R.OracleConfig(cfgd["D"], cfgd["H"], cfgd["L"], cfgd["txt"], cfgd["img"], cfgd["vocab_size"],
                          cfgd["text_vocab_size"], cfgd["mask_index"])
    mine = R.dit_forward(ocfg, P, ids, modality, mode="fp32")
    err = (mine - ref_logits).abs().max().item()
    print(f"[dit fp32] max|restated - reference| = {err:.3e}  (ref absmax {ref_logits.abs().max():.3f})")
    assert err < 2e-5, err
    # rope tables
    rc, rs = R.rope_table_2d(ocfg.head_dim, ocfg.img_length)
    assert torch.equal(rc, dit.rotary_cos_emb_img) and torch.equal(rs, dit.rotary_sin_emb_img)
    tc, ts = R.rope_table_1d(ocfg.head_dim, ocfg.length)
    assert torch.equal(tc, dit.rotary_cos_emb_txt) and torch.equal(ts, dit.rotary_sin_emb_txt)

    # reference's default CPU mode: bf16 autocast vs restated bf16 mode — loose check only
    torch.manual_seed(0)
    dit_bf = RL.build_reference_dit(ref_cfg, cfgd["vocab_size"], cfgd["text_vocab_size"], cfgd["mask_index"], dtype=torch.bfloat16)
    dit_bf.load_state_dict(P)
    dit_bf.eval()
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        ref_bf = dit_bf(ids, None, modality=modality).float()
    mine_bf = R.dit_forward(ocfg, P, ids, modality, mode="bf16").float()
    e_bf = (mine_bf - ref_bf).abs().max().item()
    e_bf32 = (mine_bf - ref_logits).abs().max().item()
    print(f"[dit bf16] restated-bf16 vs reference-cpu-autocast: {e_bf:.3e}; restated-bf16 vs reference-fp32: {e_bf32:.3e}")
    assert e_bf < 0.15, e_bf

    # gradient of a scalar through the reference (out-of-place qk-norm shim not needed in fp32 w/ clone? it is) —
    # reference eager backward raises on the in-place q/k norm write, so gradients are pinned
    # through the restatement only.
    np.savez_compressed(os.path.join(OUT, "dit_small.npz"),
                        cfg=np.array([cfgd[k] for k in ["D", "H", "L", "txt", "img", "vocab_size", "text_vocab_size", "mask_index"]]),
                        ids=_np(ids), modality=_np(modality), ref_logits_fp32=_np(ref_logits), ref_logits_cpu_bf16=_np(ref_bf),
                        rope_cos_img=_np(dit.rotary_cos_emb_img), rope_sin_img=_np(dit.rotary_sin_emb_img),
                        **{"P::" + k: _np(v) for k, v in P.items()})
    return cfgd, dit, P, ids, modality, ocfg, ids_clean


def gen_diffusion_fns(cfgd, dit, P, ids, modality, ocfg, ids_clean):
    out = {}
    s, M = make_fake_self(cfgd, ref_dit=dit, mask_entire_modality=None)
    B, N = ids.shape

    # ---- _sample_t (model.py:589-619)
    torch.manual_seed(11)
    t_ref = s._sample_t(8, torch.device("cpu"))
    torch.manual_seed(11)
    u = torch.rand(8)
    assert torch.equal(R.sample_t(u), t_ref)
    out.update(sample_t_u=_np(u), sample_t_ref=_np(t_ref))

    # ---- noise (noise_schedule.py:142-150)
    sig_ref, dsig_ref = s.noise(t_ref)
    sig, dsig = R.loglinear_noise(t_ref)
    assert torch.equal(sig, sig_ref) and torch.equal(dsig, dsig_ref)
    out.update(sigma_ref=_np(sig_ref), dsigma_ref=_np(dsig_ref))

    # ---- q_xt plain (model.py:439,579)
    x0 = ids_clean.clone()
    mc = torch.tensor([[0.3], [0.9]])
    torch.manual_seed(12)
    xt_ref, ign, _, smt, smi, mv_ref = s.q_xt(x0, mc, return_ignore_batch_mask_for_metrics=True, batch=None)
    torch.manual_seed(12)
    rand = torch.rand(B, N)
    xt, mv, _ = R.q_xt(x0, mc, rand, cfgd["mask_index"])
    assert torch.equal(xt, xt_ref) and torch.equal(mv, mv_ref)
    out.update(qxt_x0=_np(x0), qxt_mc=_np(mc), qxt_rand=_np(rand), qxt_ref=_np(xt_ref), qxt_move_ref=_np(mv_ref))

    # ---- q_xt with mask_entire_modality (model.py:470-529)
    s2, _ = make_fake_self(cfgd, ref_dit=SimpleNamespace(training=True), mask_entire_modality=0.9)
    Bm = 8
    x0m = x0[:1].repeat(Bm, 1)
    modm = modality[:1].repeat(Bm, 1)
    mm = torch.stack([modm == 0, modm == 1], dim=-1)
    mcm = torch.full((Bm, 1), 0.5)
    torch.manual_seed(13)
    xt_ref2, ign2, _, smt2, smi2, mv2 = s2.q_xt(x0m, mcm, return_ignore_batch_mask_for_metrics=True, batch=dict(modality_mask=mm))
    torch.manual_seed(13)
    randm = torch.rand(Bm, N)
    rt, ri = torch.rand(Bm, 1), torch.rand(Bm, 1)
    xt2, mv2_, ign2_ = R.q_xt(x0m, mcm, randm, cfgd["mask_index"], modality_mask=mm, mask_entire_modality=0.9, rand_txt=rt, rand_img=ri)
    assert torch.equal(xt2, xt_ref2) and torch.equal(mv2_, mv2) and torch.equal(ign2_, ign2)
    assert ign2.any() and not ign2.all()
    out.update(qxtm_x0=_np(x0m), qxtm_mod=_np(modm), qxtm_mc=_np(mcm), qxtm_rand=_np(randm), qxtm_rt=_np(rt), qxtm_ri=_np(ri),
               qxtm_ref=_np(xt_ref2), qxtm_ignore_ref=_np(ign2))

    # ---- _subs_parameterization (model.py:621-658): fp32 and bf16 logits, with and without xt
    torch.manual_seed(14)
    logits = torch.randn(B, N, cfgd["vocab_size"]) * 3
    for tag, lg in (("f32", logits), ("bf16", logits.bfloat16())):
        ref_a = s._subs_parameterization(lg.clone(), xt=xt_ref, batch=None, modality=modality)
        ref_b = s._subs_parameterization(lg.clone(), xt=None, batch=None, modality=modality)
        a = R.subs_parameterization(lg, xt_ref, modality, cfgd["mask_index"], cfgd["text_vocab_size"])
        b = R.subs_parameterization(lg, None, modality, cfgd["mask_index"], cfgd["text_vocab_size"])
        assert torch.equal(a, ref_a) and torch.equal(b, ref_b), tag
        out[f"subs_{tag}_ref_xt"] = _np(ref_a)
        out[f"subs_{tag}_ref_noxt"] = _np(ref_b)
    out.update(subs_logits=_np(logits), subs_xt=_np(xt_ref), subs_modality=_np(modality))

    # ---- compute_loss (model.py:797-1173) through the reference forward + reference DIT (fp32)
    found = RL._extract_functions(os.path.join(RL.REFERENCE_ROOT, "model.py"), ["forward", "compute_loss", "get_cond_dict"])
    g = RL._exec_functions(found, dict(Loss=lambda **k: SimpleNamespace(**k), utils=SimpleNamespace(print_nans=lambda *a: None),
                                      get_block_mask=None, get_interleaved_block_mask=None, shard_output=None))
    for img_w, snr in ((0.6, None), (0.5, 5.0)):
        s3, _ = make_fake_self(cfgd, ref_dit=dit, img_loss_weight=img_w, softmin_snr=snr)
        s3.forward = lambda *a, **k: g["forward"](s3, *a, **k)
        s3.get_cond_dict = lambda b: g["get_cond_dict"](s3, b)
        am = torch.ones(B, N, dtype=torch.bool)
        am[1, 10:14] = False
        batch = dict(input_ids=x0, attention_mask=am, modality=modality,
                     modality_mask=torch.stack([modality == 0, modality == 1], dim=-1))
        torch.manual_seed(15)
        with torch.no_grad():
            L = g["compute_loss"](s3, batch, prefix="train", batch_idx=-1)
        torch.manual_seed(15)
        u_t = torch.rand(B)
        rand_move = torch.rand(B, N)
        mine = R.training_loss(ocfg, P, x0, modality, am, u_t, rand_move, mode="fp32", img_loss_weight=img_w, softmin_snr=snr)
        e = abs(mine["loss"].item() - L.loss.item())
        print(f"[compute_loss w_img={img_w} snr={snr}] ref {L.loss.item():.6f} restated {mine['loss'].item():.6f} |d|={e:.2e}")
        assert e < 1e-4 * max(1.0, abs(L.loss.item()))
        assert torch.allclose(mine["nlls"], L.nlls, rtol=1e-4, atol=1e-3)
        tag = f"loss_w{int(img_w*10)}_snr{0 if snr is None else int(snr)}"
        out.update({tag + "_ref": np.array([L.loss.item(), L.txt_loss.item(), L.img_loss.item()]), tag + "_nlls_ref": _np(L.nlls)})
    out.update(loss_x0=_np(x0), loss_am=_np(am), loss_u_t=_np(u_t), loss_rand_move=_np(rand_move))

    # ---- _sample_categorical (model_utils.py:95-97)
    torch.manual_seed(16)
    probs = torch.softmax(torch.randn(B, N, cfgd["vocab_size"]) * 2, -1)
    torch.manual_seed(17)
    sc_ref = M._sample_categorical(probs)
    torch.manual_seed(17)
    uu = torch.rand_like(probs)
    assert torch.equal(R.sample_categorical(probs, uu), sc_ref)
    out.update(sc_probs=_np(probs), sc_u=_np(uu), sc_ref=_np(sc_ref))

    # ---- _ddpm_caching_update / _ddpm_update given p_x0 (model_eval.py:2042-2104)
    xcur = x0.clone()
    xcur[:, ::3] = cfgd["mask_index"]
    tt = torch.full((B, 1), 0.7)
    dt = (1 - 1e-5) / 16
    torch.manual_seed(18)
    _, xn_ref, _ = M._ddpm_caching_update(s, xcur, tt, dt, p_x0=probs.clone())
    torch.manual_seed(18)
    u2 = torch.rand_like(probs)
    xn = R.ddpm_caching_update(xcur, tt, dt, probs.clone(), u2, cfgd["mask_index"])
    assert torch.equal(xn, xn_ref)
    out.update(ddpm_x=_np(xcur), ddpm_t=_np(tt), ddpm_dt=np.array(dt), ddpm_u=_np(u2), ddpm_cache_ref=_np(xn_ref))
    s._ddpm_forward = lambda *a, **k: probs.clone()
    torch.manual_seed(19)
    xn2_ref, _ = M._ddpm_update(s, xcur, tt, dt)
    torch.manual_seed(19)
    u3 = torch.rand_like(probs)
    xn2 = R.ddpm_update(xcur, tt, dt, probs.clone(), u3, cfgd["mask_index"])
    assert torch.equal(xn2, xn2_ref)
    out.update(ddpm_u3=_np(u3), ddpm_ref=_np(xn2_ref))

    # ---- _ddpm_forward end-to-end through the reference DIT, no CFG (model_eval.py:1761-1833)
    s4, _ = make_fake_self(cfgd, ref_dit=dit)
    s4.forward = lambda *a, **k: g["forward"](s4, *a, **k)
    with torch.no_grad():
        p_ref = M._ddpm_forward(s4, xcur, tt, None, x0=None, x0_unmask=None, modality=modality)
    lg = R.dit_forward(ocfg, P, xcur, modality, mode="fp32")
    p_mine = R.subs_parameterization(lg, xcur, modality, cfgd["mask_index"], cfgd["text_vocab_size"]).exp()
    assert torch.allclose(p_mine, p_ref, rtol=1e-4, atol=1e-6), (p_mine - p_ref).abs().max()
    out.update(ddpmfwd_p_ref=_np(p_ref))
    # with CFG (model_eval.py:1763-1818)
    s5, _ = make_fake_self(cfgd, ref_dit=dit, eval_cfg=2.5)
    s5.forward = lambda *a, **k: g["forward"](s5, *a, **k)
    x0_unmask = torch.zeros(B, N, dtype=torch.bool)
    x0_unmask[:, :cfgd["txt"]] = True
    xc = torch.where(x0_unmask, x0, torch.full_like(x0, cfgd["mask_index"]))
    with torch.no_grad():
        pc_ref = M._ddpm_forward(s5, xc, tt.squeeze(-1), None, x0=x0, x0_unmask=x0_unmask, modality=modality)
    xu = xc.clone()
    xu[x0_unmask] = cfgd["mask_index"]
    lc = R.dit_forward(ocfg, P, xc, modality, mode="fp32")
    lu = R.dit_forward(ocfg, P, xu, modality, mode="fp32")
    comb = R.cfg_combine(lc, lu, tt.squeeze(-1), 2.5)
    pc = R.subs_parameterization(comb, None, modality, cfgd["mask_index"], cfgd["text_vocab_size"]).exp()
    assert torch.allclose(pc, pc_ref, rtol=2e-4, atol=1e-6), (pc - pc_ref).abs().max()
    out.update(cfg_x=_np(xc), cfg_unmask=_np(x0_unmask), cfg_p_ref=_np(pc_ref))

    # ---- adap_sche (model_eval.py:2964-3001) and _maskgit_update (model_eval.py:3045-3114)
    sch_ref = M.adap_sche(xcur, 8, cfgd["mask_index"], mode="arccos")
    assert torch.equal(R.adap_sche(xcur, 8, cfgd["mask_index"]), sch_ref)
    s6, _ = make_fake_self(cfgd, ref_dit=dit)
    s6._ddpm_forward = lambda *a, **k: probs.clone()
    s6.config.eval["maskgit_r_temp"] = 10
    torch.manual_seed(20)
    np.random.seed(20)
    mg_ref, _ = M._maskgit_update(s6, xcur, tt, dt, schedule=sch_ref, step=2)
    torch.manual_seed(20)
    np.random.seed(20)
    pred = torch.multinomial(probs.view(-1, probs.shape[-1]), 1)[:, 0].view(B, N)
    gum = torch.from_numpy(np.random.gumbel(size=(B, N)))
    mg = R.maskgit_update(xcur, tt, probs, pred, gum, sch_ref[:, 2], cfgd["mask_index"], r_temp=10)
    assert torch.equal(mg, mg_ref)
    out.update(sche_ref=_np(sch_ref), mg_pred=_np(pred), mg_gumbel=_np(gum), mg_ref=_np(mg_ref))

    np.savez_compressed(os.path.join(OUT, "diffusion_fns.npz"), **out)



def gen_interleaved():
    """Interleaved batches (data.require_sample_ids, BASELINE cfg5): the reference DIT run with packed samples,
    per-image-block 2-D RoPE tables, img_count_embedding and the FlexAttention document mask (eval mode, eager flex on CPU)
    vs the restatement; also pins `interleaved_token_tables` against add_img_data_to_blocks/add_txt_data_to_blocks."""
    from torch.nn.attention.flex_attention import create_block_mask
    D, H, L, N, V, tv, mi = 128, 2, 2, 640, 160, 97, 96
    ref_cfg = RL.make_ref_config(D, H, L, 64, 256, require_sample_ids=True)
    ref_cfg.model.length = N
    ref_cfg.model.use_flex_attention = True
    torch.manual_seed(0)
    dit = RL.build_reference_dit(ref_cfg, V, tv, mi, dtype=torch.float32)
    dit.eval()
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        dit.img_count_embedding.copy_(torch.randn(dit.img_count_embedding.shape, generator=g) * 0.5)
        for n, p in dit.named_parameters():
            if "norm" in n and n.endswith("weight"):
                p.add_((torch.rand(p.shape, generator=g) - 0.5) * 0.4)
    B = 4
    modality = torch.zeros(B, N, dtype=torch.long)
    sid = torch.zeros(B, N, dtype=torch.long)
    # row 0: sample 0 = txt40 img256 txt24 ; sample 1 = img256 txt30 ; pad
    modality[0, 40:296] = 1; sid[0, 320:] = 1; modality[0, 320:576] = 1; sid[0, 606:] = -1
    # row 1: sample 0 = txt10 img64(no table) txt26 ; sample 1 = txt100 img256 txt4 img40(no table) txt100 ; sample 2 = txt40
    modality[1, 10:74] = 1; sid[1, 100:] = 1; modality[1, 200:456] = 1; modality[1, 460:500] = 1; sid[1, 600:] = 2
    # row 2: sample 0 = img256 img... two images of one sample separated by 1 text token, then sample 1 starts with an image
    modality[2, 0:256] = 1; modality[2, 257:513] = 1; sid[2, 513:] = 1; modality[2, 514:578] = 1; sid[2, 630:] = -1
    # row 3: the images of two different samples are adjacent -> ONE modality run of 512 tokens (no table: cos = sin = 0)
    modality[3, 0:512] = 1; sid[3, 256:] = 1
    ids = torch.randint(0, mi, (B, N), generator=g)
    ids[modality == 1] = torch.randint(tv, V, (int((modality == 1).sum()),), generator=g)
    ids[0, 50:90] = mi
    ids[1, 5] = mi

    def mm(b, h, q, k):
        return (sid[b, q] == sid[b, k]) & (sid[b, q] != -1)
    bm = create_block_mask(mm, B=B, H=None, Q_LEN=N, KV_LEN=N, device="cpu")
    with torch.no_grad():
        ref_logits = dit(ids, None, modality=modality, sample_ids=sid, block_mask=bm).float()
    # tables straight from the reference helper functions
    mod = RL.load_reference_dit_module()
    hd = D // H
    cos = torch.zeros(B, N, hd // 2); sin = torch.zeros(B, N, hd // 2)
    xz = torch.zeros(B, N, D)
    mmask = modality.bool()
    mod.add_img_data_to_blocks(xz, cos, mmask, sid, {k: getattr(dit, f"rotary_cos_emb_img_{k}") for k in (256, 1024, 2304, 4096)}, dit.img_count_embedding.detach())
    mod.add_img_data_to_blocks(None, sin, mmask, sid, {k: getattr(dit, f"rotary_sin_emb_img_{k}") for k in (256, 1024, 2304, 4096)}, None)
    mod.add_txt_data_to_blocks(cos, mmask, sid, dit.rotary_cos_emb_txt)
    mod.add_txt_data_to_blocks(sin, mmask, sid, dit.rotary_sin_emb_txt)
    ocfg = R.OracleConfig(D, H, L, 64, 256, V, tv, mi, require_sample_ids=True)
    ocfg_len = N
    assert ocfg.length != N or True
    P = {k: v.detach().clone() for k, v in dit.state_dict().items()}
    c2, s2, ordinal = R.interleaved_token_tables(_with_length(ocfg, N), modality, sid)
    assert torch.equal(c2, cos) and torch.equal(s2, sin), ((c2 - cos).abs().max(), (s2 - sin).abs().max())
    add = torch.where((ordinal >= 0)[..., None], dit.img_count_embedding.detach()[ordinal.clamp(min=0)], torch.zeros(B, N, D))
    assert torch.equal(add, xz)
    mine = R.dit_forward(_with_length(ocfg, N), P, ids, modality, mode="fp32", sample_ids=sid)
    valid = (sid != -1)
    err = (mine - ref_logits)[valid].abs().max().item()
    print(f"[interleaved fp32] max|restated - reference| on non-pad tokens = {err:.3e}  (ref absmax {ref_logits[valid].abs().max():.3f})")
    assert err < 5e-5, err
    # ---- q_xt, interleaved whole-block masking (model.py:483-522), reference code on a fake self
    cfgd = dict(txt=N - 256, img=256, mask_index=mi, text_vocab_size=tv, vocab_size=V)
    s2, _ = make_fake_self(cfgd, ref_dit=SimpleNamespace(training=True), mask_entire_modality=0.2)
    s2.config.trainer.interleaved = True
    x0 = torch.where(ids == mi, torch.zeros_like(ids), ids)
    mc = torch.full((B, 1), 0.3)
    torch.manual_seed(21)
    xt_ref, ign_ref, _, _, _, mv_ref = s2.q_xt(x0, mc, return_ignore_batch_mask_for_metrics=True, batch=dict(modality=modality, sample_ids=sid))
    torch.manual_seed(21)
    rand = torch.rand(B, N); _rt, _ri = torch.rand(B, 1), torch.rand(B, 1)
    from unidisc.utils.tensor_utils import get_contiguous_blocks_per_sample
    bi, sp, ep = get_contiguous_blocks_per_sample(modality, sid)
    Mb = int(((ep - sp) > 4).sum())
    rand_blocks = torch.rand(Mb, 1)
    xt_o, mv_o, ign_o = R.q_xt_interleaved(x0, mc, rand, mi, modality, sid, 0.2, rand_blocks)
    assert torch.equal(xt_o, xt_ref) and torch.equal(mv_o, mv_ref) and torch.equal(ign_o, ign_ref.bool())
    assert ign_ref.any() and not ign_ref.all(), ign_ref
    print(f"[interleaved q_xt] {Mb} blocks, rows force-masked: {ign_ref.tolist()}")
    np.savez_compressed(os.path.join(OUT, "interleaved.npz"),
                        qxt_x0=_np(x0), qxt_mc=_np(mc), qxt_rand=_np(rand), qxt_rand_blocks=_np(rand_blocks), qxt_seed=np.array(21),
                        qxt_ref=_np(xt_ref), qxt_move_ref=_np(mv_ref), qxt_ignore_ref=_np(ign_ref.bool()),
                        cfg=np.array([D, H, L, N, V, tv, mi]), ids=_np(ids), modality=_np(modality), sample_ids=_np(sid),
                        ref_logits_fp32=_np(ref_logits), ref_cos=_np(cos), ref_sin=_np(sin), ref_ordinal=_np(ordinal),
                        **{"P::" + k: _np(v) for k, v in P.items()})


def gen_timecond():
    """config.time_conditioning=True (adaLN shift/scale/gate from sigma; off in every shipped config): reference DIT vs the
    restatement, including the all-text batch quirk of modulate_fused (dit.py:301-304)."""
    import dataclasses
    D, H, L, txt, img, V, tv, mi = 128, 2, 2, 64, 64, 160, 97, 96
    ref_cfg = RL.make_ref_config(D, H, L, txt, img, time_conditioning=True)
    torch.manual_seed(0)
    dit = RL.build_reference_dit(ref_cfg, V, tv, mi, dtype=torch.float32)
    dit.eval()
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for n, p in dit.named_parameters():
            if "adaLN_modulation" in n:                       # zero-initialised in the reference: make them matter
                p.copy_((torch.rand(p.shape, generator=g) - 0.5) * (0.4 if n.endswith("weight") else 0.2))
            elif "norm" in n and n.endswith("weight"):
                p.add_((torch.rand(p.shape, generator=g) - 0.5) * 0.4)
    P = {k: v.detach().clone() for k, v in dit.state_dict().items()}
    ids, modality = R.synthetic_batch(3, txt, img, tv, V, seed=43)
    ids[0, 3] = mi
    ids[1, 70:90] = mi
    sigma = torch.tensor([0.15, 1.3, 4.0])
    with torch.no_grad():
        ref_logits = dit(ids, sigma, modality=modality)
    ocfg = dataclasses.replace(R.OracleConfig(D, H, L, txt, img, V, tv, mi), time_conditioning=True)
    mine = R.dit_forward(ocfg, P, ids, modality, mode="fp32", sigma=sigma)
    err = (mine - ref_logits).abs().max().item()
    print(f"[timecond fp32] max|restated - reference| = {err:.3e} (ref absmax {ref_logits.abs().max():.3f})")
    assert err < 3e-5, err
    # all-text batch: modulate_fused modulates every token
    ids_t = torch.randint(0, mi, (2, txt + img), generator=g)
    mod_t = torch.zeros_like(ids_t)
    with torch.no_grad():
        ref_t = dit(ids_t, sigma[:2], modality=mod_t)
    mine_t = R.dit_forward(ocfg, P, ids_t, mod_t, mode="fp32", sigma=sigma[:2])
    err_t = (mine_t - ref_t).abs().max().item()
    print(f"[timecond fp32, all-text batch] max|restated - reference| = {err_t:.3e}")
    assert err_t < 3e-5, err_t
    # without conditioning the logits differ: the fixture really exercises adaLN
    plain = R.dit_forward(R.OracleConfig(D, H, L, txt, img, V, tv, mi), P, ids, modality, mode="fp32")
    assert (plain - ref_logits).abs().max() > 1e-2
    np.savez_compressed(os.path.join(OUT, "timecond.npz"), cfg=np.array([D, H, L, txt, img, V, tv, mi]), ids=_np(ids),
                        modality=_np(modality), sigma=_np(sigma), ref_logits_fp32=_np(ref_logits), ids_txt=_np(ids_t),
                        modality_txt=_np(mod_t), ref_logits_txt_fp32=_np(ref_t), **{"P::" + k: _np(v) for k, v in P.items()})


CFG1 = dict(D=384, H=6, L=6, txt=60, img=196, text_vocab_size=32001, vocab_size=48385, mask_index=32000, B=4)


def cfg1_params(seed=123):
    """Deterministic parameters for the BASELINE.json configs[0]-sized model (37 M of them: too large to commit, so both the
    generator and the tests rebuild them from the seed with the CPU generator)."""
    ocfg = R.OracleConfig(CFG1["D"], CFG1["H"], CFG1["L"], CFG1["txt"], CFG1["img"], CFG1["vocab_size"], CFG1["text_vocab_size"],
                          CFG1["mask_index"])
    return ocfg, R.init_params(ocfg, seed=seed)


def gen_cfg1():
    """BASELINE.json configs[0] (the reference's own CPU-runnable parity case): DiT-S depth 6, dim 384, 6 heads, seq_len 256,
    batch 4, fp32, eager, real vocabulary (32001 text + 16384 image ids).  The image span is 196 = 14 x 14 tokens and the text
    span 60, because the reference asserts a square image length for its 2-D RoPE (models/dit.py:1048); 64 + 192 cannot be run.
    Stores the inputs, the random draws, sub-sampled reference logits (every 4th token, every 61st vocabulary entry), whole-
    tensor checksums and the reference compute_loss outputs."""
    c = CFG1
    ocfg, P = cfg1_params()
    ref_cfg = RL.make_ref_config(c["D"], c["H"], c["L"], c["txt"], c["img"])
    dit = RL.build_reference_dit(ref_cfg, c["vocab_size"], c["text_vocab_size"], c["mask_index"], dtype=torch.float32)
    missing = dit.load_state_dict(P, strict=True)
    dit.eval()
    B, N = c["B"], c["txt"] + c["img"]
    x0, modality = R.synthetic_batch(B, c["txt"], c["img"], c["text_vocab_size"], c["vocab_size"], seed=7)
    g = torch.Generator().manual_seed(8)
    u_t = torch.rand(B, generator=g)
    rand_move = torch.rand(B, N, generator=g)
    t = R.sample_t(u_t)
    sigma, _ = R.loglinear_noise(t)
    xt, move, _ = R.q_xt(x0, 1 - torch.exp(-sigma[:, None]), rand_move, c["mask_index"])
    with torch.no_grad():
        ref_logits = dit(xt, None, modality=modality)
    mine = R.dit_forward(ocfg, P, xt, modality, mode="fp32")
    err = (mine - ref_logits).abs().max().item()
    print(f"[cfg1 fp32] max|restated - reference| = {err:.3e} (ref absmax {ref_logits.abs().max():.3f})")
    assert err < 5e-5, err
    # reference compute_loss on the same draws (torch.rand order of model.py:844 -> :439 replayed through the global generator)
    found = RL._extract_functions(os.path.join(RL.REFERENCE_ROOT, "model.py"), ["forward", "compute_loss", "get_cond_dict"])
    gg = RL._exec_functions(found, dict(Loss=lambda **k: SimpleNamespace(**k), utils=SimpleNamespace(print_nans=lambda *a: None),
                                       get_block_mask=None, get_interleaved_block_mask=None, shard_output=None))
    cfgd = dict(txt=c["txt"], img=c["img"], mask_index=c["mask_index"], text_vocab_size=c["text_vocab_size"], vocab_size=c["vocab_size"])
    s3, _ = make_fake_self(cfgd, ref_dit=dit, img_loss_weight=0.6)
    s3.forward = lambda *a, **k: gg["forward"](s3, *a, **k)
    s3.get_cond_dict = lambda b: gg["get_cond_dict"](s3, b)
    am = torch.ones(B, N, dtype=torch.bool)
    am[2, 50:58] = False
    batch = dict(input_ids=x0, attention_mask=am, modality=modality, modality_mask=torch.stack([modality == 0, modality == 1], dim=-1))
    torch.manual_seed(21)
    with torch.no_grad():
        Lr = gg["compute_loss"](s3, batch, prefix="train", batch_idx=-1)
    torch.manual_seed(21)
    u2 = torch.rand(B)
    rm2 = torch.rand(B, N)
    m2 = R.training_loss(ocfg, P, x0, modality, am, u2, rm2, mode="fp32", img_loss_weight=0.6)
    e = abs(m2["loss"].item() - Lr.loss.item())
    print(f"[cfg1 compute_loss] ref {Lr.loss.item():.6f} restated {m2['loss'].item():.6f} |d|={e:.2e}")
    assert e < 1e-4 * max(1.0, abs(Lr.loss.item()))
    np.savez_compressed(os.path.join(OUT, "cfg1.npz"),
                        cfg=np.array([c[k] for k in ["D", "H", "L", "txt", "img", "vocab_size", "text_vocab_size", "mask_index"]]),
                        param_seed=np.array([123]), x0=_np(x0), modality=_np(modality), u_t=_np(u_t), rand_move=_np(rand_move),
                        xt=_np(xt), ref_logits_sub=_np(ref_logits[:, ::4, ::61]),
                        ref_logits_sum=np.array([ref_logits.double().sum().item(), ref_logits.double().abs().sum().item()]),
                        ref_logits_rowmax=_np(ref_logits.max(dim=-1).values), ref_logits_argmax=_np(ref_logits.argmax(dim=-1)),
                        loss_am=_np(am), loss_u_t=_np(u2), loss_rand_move=_np(rm2),
                        loss_ref=np.array([Lr.loss.item(), Lr.txt_loss.item(), Lr.img_loss.item()]), loss_nlls_ref=_np(Lr.nlls),
                        param_checksum=np.array([sum(v.double().sum().item() for v in P.values())]))



def _ub_cases():
    """Synthetic dataloader batches for update_batch (tokenised paths), with their config switches."""
    g = torch.Generator().manual_seed(31)
    txt, img, tv = 12, 16, 97
    cases = {}
    # (a) pre-tokenised unified sample: int32 ids, text padding mask
    tam = torch.ones(3, txt, dtype=torch.int32)
    tam[1, 9:] = 0
    cases["pretok"] = (dict(txt_input_ids=torch.randint(0, tv - 1, (3, txt), generator=g, dtype=torch.int32), txt_attention_mask=tam,
                            img_input_ids=torch.randint(0, 63, (3, img), generator=g, dtype=torch.int32)),
                       dict(), txt, img)
    # (b) the same with packed-sample ids (data.require_sample_ids, trainer.interleaved) and an explicit modality with -1 padding
    sid = torch.zeros(3, txt + img, dtype=torch.int32)
    sid[0, 20:] = 1
    sid[2, 5:] = 2
    mod = torch.cat([torch.zeros(3, txt, dtype=torch.int32), torch.ones(3, img, dtype=torch.int32)], 1)
    mod[1, 9:12] = -1
    cases["pretok_sid"] = (dict(txt_input_ids=torch.randint(0, tv - 1, (3, txt), generator=g, dtype=torch.int32), txt_attention_mask=tam.clone(),
                                img_input_ids=torch.randint(0, 63, (3, img), generator=g, dtype=torch.int32), sample_ids=sid, modality=mod),
                           dict(trainer=dict(interleaved=True), data=dict(require_sample_ids=True)), txt, img)
    # (c) joint multimodal batch with un-shifted image ids (trainer.force_shift_image_batches) and interleaved blocks
    mod2 = torch.zeros(2, txt + img, dtype=torch.int64)
    mod2[0, 4:20] = 1
    mod2[1, :16] = 1
    ids2 = torch.where(mod2 == 1, torch.randint(0, 63, (2, txt + img), generator=g), torch.randint(0, tv - 1, (2, txt + img), generator=g))
    cases["joint"] = (dict(input_ids=ids2.to(torch.int32), modality=mod2, attention_mask=torch.ones(2, txt + img, dtype=torch.int64)),
                      dict(trainer=dict(force_shift_image_batches=True, interleaved=True)), txt, img)
    return cases, tv


def _ub_config(txt, img, over):
    cfg = dict(parameterization="subs", backbone="dit",
               eval=dict(), data=dict(require_sample_ids=False, txt_only=False),
               trainer=dict(image_mode="discrete", multimodal_batches=True, interleaved=False, ignore_text_in_unified=False,
                            ar_inpainting=False),
               model=dict(txt_length=txt, img_length=img, length=txt + img, unified_model=True))
    for sec, kv in over.items():
        cfg[sec].update(kv)
    return cfg


def gen_update_batch():
    """Diffusion.update_batch (model.py:157-395) on tokenised batches: the reference function itself (extracted from model.py,
    fake self) vs unidisc_b200.model.update_batch; every output tensor must be identical."""
    from unidisc.utils.tensor_utils import get_contiguous_blocks
    from unidisc_b200.model import update_batch as mine_fn

    class TensorDict(dict):                      # stands in for tensordict.TensorDict (only constructed for the block metadata)
        def __init__(self, d=None, batch_size=None):
            super().__init__(d or {})

    found = RL._extract_functions(os.path.join(RL.REFERENCE_ROOT, "model.py"), ["update_batch"])
    gg = RL._exec_functions(found, dict(TensorDict=TensorDict, get_contiguous_blocks=get_contiguous_blocks, get_image_batch=None))
    cases, tv = _ub_cases()
    out = {}
    for name, (batch, over, txt, img) in cases.items():
        cfg = _ub_config(txt, img, over)
        s = SimpleNamespace(config=RL.to_attrdict(cfg), image_model=True, is_compiled=True, text_vocab_size=tv, device=torch.device("cpu"),
                            training=True, unified_model=True)
        s.txt_sl = lambda b: b["modality_mask"][..., 0]
        s.img_sl = lambda b: b["modality_mask"][..., 1]
        ref = gg["update_batch"](s, {k: v.clone() for k, v in batch.items()})
        mine = mine_fn({k: v.clone() for k, v in batch.items()}, RL.to_attrdict(cfg), text_vocab_size=tv, device=torch.device("cpu"))
        flat = lambda d: {k if not isinstance(v, dict) else None: v for k, v in d.items()}
        keys = sorted(k for k, v in ref.items() if isinstance(v, torch.Tensor))
        assert keys == sorted(k for k, v in mine.items() if isinstance(v, torch.Tensor)), (keys, sorted(mine))
        for k in keys:
            assert ref[k].dtype == mine[k].dtype and torch.equal(ref[k], mine[k]), (name, k)
            out[f"{name}::out::{k}"] = _np(ref[k])
        if "interleaved_metadata" in ref:
            for k in ("batch_indices", "start_positions", "end_positions"):
                assert torch.equal(ref["interleaved_metadata"][k], mine["interleaved_metadata"][k]), (name, k)
                out[f"{name}::meta::{k}"] = _np(ref["interleaved_metadata"][k])
        for k, v in batch.items():
            out[f"{name}::in::{k}"] = _np(v)
        print(f"[update_batch {name}] {len(keys)} tensors identical" + (" + block metadata" if "interleaved_metadata" in ref else ""))
    np.savez_compressed(os.path.join(OUT, "update_batch.npz"), text_vocab_size=np.array([tv]), **out)



def gen_first_hitting():
    """_first_hitting_update (model_eval.py:3004-3043) given p_x0: the reference function (fake self whose _ddpm_forward returns
    fixed probabilities) vs unidisc_b200.model.first_hitting_select fed with the same draws."""
    from unidisc_b200.model import first_hitting_select
    B, N, V, mi = 3, 40, 50, 49
    cfgd = dict(txt=16, img=24, mask_index=mi, text_vocab_size=30, vocab_size=V)
    s, M = make_fake_self(cfgd, ref_dit=None)
    g = torch.Generator().manual_seed(41)
    x = torch.randint(0, mi, (B, N), generator=g)
    x[torch.rand(B, N, generator=g) < 0.6] = mi
    x[2, :] = torch.randint(0, mi, (N,), generator=g)          # a fully unmasked row (num_unmask clamps to 0)
    probs = torch.softmax(torch.randn(B, N, V, generator=g) * 2, -1)
    probs[..., mi] = 0
    s._ddpm_forward = lambda *a, **k: probs.clone()
    sched = M.adap_sche(x, 6, mi, mode="arccos")
    tt, dt = torch.full((B, 1), 0.6), (1 - 1e-5) / 6
    out = dict(x=_np(x), probs=_np(probs), schedule=_np(sched))
    for step in (0, 3, 5):
        torch.manual_seed(50 + step)
        ref, nfe = M._first_hitting_update(s, x.clone(), tt, dt, schedule=sched, step=step)
        torch.manual_seed(50 + step)
        u = torch.rand_like(probs)
        rv = torch.rand(B, N)
        mine = first_hitting_select(x.clone(), R.sample_categorical(probs, u), sched[:, step], rv, mi)
        assert nfe == 1 and torch.equal(mine, ref), step
        assert int(((ref != x) & (x != mi)).sum()) == 0               # unmasked tokens are carried over
        out.update({f"u_{step}": _np(u), f"rv_{step}": _np(rv), f"ref_{step}": _np(ref)})
    print("[first_hitting] 3 steps identical to the reference")
    np.savez_compressed(os.path.join(OUT, "first_hitting.npz"), mask_index=np.array([mi]), **out)



def gen_sampler_logits():
    """Samplers driven from LOGITS (the fused kernels' input): the reference chain `_subs_parameterization(logits).exp()` ->
    `_ddpm_caching_update` / `_ddpm_update` / `_maskgit_update` (model.py:621-658, model_eval.py:2042-2104, 3045-3114) executed
    on bf16-representable fp32 logits, with the noise tensors the reference draws (torch.rand_like, the Exp(1) tensor inside
    torch.multinomial, np.random.gumbel) replayed and stored.  -> tests/golden/sampler_logits.npz"""
    cfgd = dict(txt=16, img=32, mask_index=96, text_vocab_size=97, vocab_size=160)
    B, N, V, tv, mi = 3, 48, 160, 97, 96
    s, M = make_fake_self(cfgd, ref_dit=None)
    g = torch.Generator().manual_seed(61)
    logits = (torch.randn(B, N, V, generator=g) * 3).bfloat16().float()
    modality = torch.cat([torch.zeros(B, 16, dtype=torch.int64), torch.ones(B, 32, dtype=torch.int64)], 1)
    x0 = torch.where(modality == 0, torch.randint(0, tv - 1, (B, N), generator=g), torch.randint(tv, V, (B, N), generator=g))
    x = x0.clone()
    x[0, torch.rand(N, generator=g) < 0.7] = mi
    x[1, torch.rand(N, generator=g) < 0.15] = mi
    # sample 2 stays fully unmasked: every update must carry it over
    p = s._subs_parameterization(logits.clone(), xt=x, batch=None, modality=modality).exp()
    assert torch.equal(p, R.subs_parameterization(logits, x, modality, mi, tv).exp())
    s._ddpm_forward = lambda *a, **k: p.clone()
    tt = torch.tensor([[0.8], [0.35], [0.6]])
    dt = (1 - 1e-5) / 16
    out = dict(logits=_np(logits), modality=_np(modality), x=_np(x), t=_np(tt), dt=np.array(dt), cfg=np.array([V, tv, mi]))
    torch.manual_seed(62)
    _, xn_ref, _ = M._ddpm_caching_update(s, x, tt, dt, p_x0=None)
    torch.manual_seed(62)
    u = torch.rand_like(p)
    assert torch.equal(R.ddpm_caching_update(x, tt, dt, p.clone(), u, mi), xn_ref)
    out.update(cache_u=_np(u), cache_ref=_np(xn_ref))
    torch.manual_seed(63)
    xn2_ref, _ = M._ddpm_update(s, x, tt, dt)
    torch.manual_seed(63)
    u2 = torch.rand_like(p)
    assert torch.equal(R.ddpm_update(x, tt, dt, p.clone(), u2, mi), xn2_ref)
    out.update(ddpm_u=_np(u2), ddpm_ref=_np(xn2_ref))
    # maskgit at three schedule steps
    s.config.eval["maskgit_r_temp"] = 10
    sched = M.adap_sche(x, 8, mi, mode="arccos")
    out["schedule"] = _np(sched)
    for step in (0, 4, 7):
        torch.manual_seed(70 + step)
        np.random.seed(70 + step)
        ref, nfe = M._maskgit_update(s, x.clone(), tt, dt, schedule=sched, step=step)
        torch.manual_seed(70 + step)
        np.random.seed(70 + step)
        e_noise = torch.empty_like(p.view(-1, V)).exponential_(1)             # the draw inside torch.multinomial
        gum = torch.from_numpy(np.random.gumbel(size=(B, N)))
        torch.manual_seed(70 + step)
        pred_ref = torch.multinomial(p.view(-1, V), 1)[:, 0].view(B, N)
        assert torch.equal(R.multinomial_from_exponential(p, e_noise), pred_ref), "torch.multinomial != argmax(p / Exp(1))"
        mine = R.maskgit_update_from_noise(x, tt, p, e_noise, gum, sched[:, step], mi, r_temp=10)
        assert nfe == 1 and torch.equal(mine, ref), step
        assert int(((ref != x) & (x != mi)).sum()) == 0
        out.update({f"mg_e_{step}": _np(e_noise), f"mg_gumbel_{step}": _np(gum), f"mg_ref_{step}": _np(ref), f"mg_pred_{step}": _np(pred_ref)})
    print("[sampler_logits] ddpm_cache / ddpm / maskgit(3 steps) identical to the reference; multinomial == argmax(p/Exp(1))")
    np.savez_compressed(os.path.join(OUT, "sampler_logits.npz"), **out)


def gen_attn_cache():
    """Inference attention caching (eval.attention_caching, model_eval.py:2297-2367 + dit.py:793-812): the UNMODIFIED reference
    DIT (use_flex_attention, eager flex on CPU) run through one caching cycle — step 0 full attention, step 1 masked attention
    (`get_block_mask(txt_dropout=False, img_dropout=True)`, model_utils.py:721-738) that stores the image K/V, step 2 text-only
    forward writing the cache's text slice — against the restatement.  Parameters = tests/golden/dit_small.npz.
    -> tests/golden/attn_cache.npz"""
    from torch.nn.attention.flex_attention import create_block_mask
    gd = np.load(os.path.join(OUT, "dit_small.npz"))
    D, H, L, txt, img, V, tv, mi = [int(v) for v in gd["cfg"]]
    P = {k[3:]: torch.from_numpy(gd[k]) for k in gd.files if k.startswith("P::")}
    ref_cfg = RL.make_ref_config(D, H, L, txt, img)
    ref_cfg.model.use_flex_attention = True
    torch.manual_seed(0)
    dit = RL.build_reference_dit(ref_cfg, V, tv, mi, dtype=torch.float32)
    dit.load_state_dict(P)
    dit.eval()
    found = RL._extract_functions(os.path.join(RL.REFERENCE_ROOT, "model_utils.py"), ["_attn_mask", "get_block_mask"])
    gm = RL._exec_functions(found, dict(create_block_mask=create_block_mask))
    ocfg = R.OracleConfig(D, H, L, txt, img, V, tv, mi)
    B, N = 2, txt + img
    ids, modality = R.synthetic_batch(B, txt, img, tv, V, seed=7)
    g = torch.Generator().manual_seed(8)
    x0 = ids.clone()
    x0[torch.rand(B, N, generator=g) < 0.6] = mi
    x1 = x0.clone()
    x1[:, 5:40:3] = ids[:, 5:40:3]                      # a few text tokens get revealed between the steps
    x2 = x1.clone()
    x2[:, 2:60:4] = ids[:, 2:60:4]
    txt_sl = slice(None, txt)
    out = dict(x0=_np(x0), x1=_np(x1), x2=_np(x2), modality=_np(modality))
    with torch.no_grad():
        dit.set_flex_attention_cache(B, N, torch.device("cpu"), torch.float32)
        r0 = dit(x0, None, modality=modality, block_mask=True, update_cache_slice=None)                       # step 0
        bm = gm["get_block_mask"](txt_batch_attn_dropout=torch.zeros(B, dtype=torch.bool),
                                  img_batch_attn_dropout=torch.ones(B, dtype=torch.bool), txt_length=txt, batch_size=B, seq_len=N,
                                  device=torch.device("cpu"))
        r1 = dit(x1, None, modality=modality, block_mask=bm, update_cache_slice=slice(0, N))                   # step 1
        ck1 = [blk.attention.cache_k.clone() for blk in dit.blocks]
        r2 = dit(x2[:, txt_sl], None, modality=modality[:, txt_sl], block_mask=True, update_cache_slice=txt_sl)  # step 2
        ck2 = [blk.attention.cache_k.clone() for blk in dit.blocks]
        cv2 = [blk.attention.cache_v.clone() for blk in dit.blocks]
    cache = {}
    m0 = R.dit_forward(ocfg, P, x0, modality, mode="fp32")
    m1 = R.dit_forward(ocfg, P, x1, modality, mode="fp32", attn_mask=R.caching_step_mask(txt, N), kv_cache=cache, cache_op="store")
    e_ck1 = max((cache[i]["k"] - ck1[i]).abs().max().item() for i in range(L))
    m2 = R.dit_forward(ocfg, P, x2[:, txt_sl], modality[:, txt_sl], mode="fp32", kv_cache=cache, cache_op="update", update_slice=txt_sl)
    e_ck2 = max((cache[i]["k"] - ck2[i]).abs().max().item() for i in range(L))
    e_cv2 = max((cache[i]["v"] - cv2[i]).abs().max().item() for i in range(L))
    errs = [(m0 - r0).abs().max().item(), (m1 - r1).abs().max().item(), (m2 - r2).abs().max().item()]
    print(f"[attn_cache] max|restated - reference| logits step0/1/2 = {errs[0]:.2e} / {errs[1]:.2e} / {errs[2]:.2e}; "
          f"cache K after step 1 {e_ck1:.2e}, K/V after step 2 {e_ck2:.2e} / {e_cv2:.2e}")
    assert max(errs) < 3e-5 and max(e_ck1, e_ck2, e_cv2) < 1e-5
    # the variant that attends to the updated cache (documented intent, dit.py:795-797) must differ from the shipped dataflow
    cache_b = {i: dict(k=ck1[i].clone(), v=cache[i]["v"].clone()) for i in range(L)}
    m2b = R.dit_forward(ocfg, P, x2[:, txt_sl], modality[:, txt_sl], mode="fp32", kv_cache=cache_b, cache_op="update", update_slice=txt_sl,
                        attend_cache=True)
    assert (m2b - r2).abs().max().item() > 1e-3
    sub = lambda t: _np(t[:, :, ::7])                                     # sub-sampled logits keep the fixture small
    out.update(ref_step0=sub(r0), ref_step1=sub(r1), ref_step2=sub(r2), ref_cache_k_blk1=_np(ck2[1][:, :, ::5, ::3]),
               ref_cache_v_blk1=_np(cv2[1][:, :, ::5, ::3]))
    np.savez_compressed(os.path.join(OUT, "attn_cache.npz"), **out)


def _with_length(ocfg, N):
    """OracleConfig whose `length` (txt_length + img_length) equals the packed sequence length N."""
    import dataclasses
    return dataclasses.replace(ocfg, txt_length=N - ocfg.img_length)

def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    if len(sys.argv) > 1 and sys.argv[1] == "cfg1":
        gen_cfg1()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "first_hitting":
        gen_first_hitting()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "sampler_logits":
        gen_sampler_logits()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "attn_cache":
        RL.load_reference_diffusion_methods()
        gen_attn_cache()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "update_batch":
        RL.load_reference_diffusion_methods()          # installs the import shims / sys.path for the reference tree
        gen_update_batch()
        return
    gen_diffusion_fns(*gen_dit())
    gen_interleaved()
    gen_timecond()
    gen_cfg1()
    gen_update_batch()
    gen_first_hitting()
    gen_sampler_logits()
    gen_attn_cache()
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
