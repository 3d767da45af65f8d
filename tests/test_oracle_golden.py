"""CPU: the restated oracle (oracle/restated.py) against the committed golden vectors that
oracle/gen_golden.py produced by running the unmodified reference code (tests/golden/*.npz)."""
import numpy as np
import pytest
import torch

from oracle import restated as R


def _cfg(g):
    D, H, L, txt, img, V, tv, mi = [int(v) for v in g["cfg"]]
    return R.OracleConfig(D, H, L, txt, img, V, tv, mi)


def _params(g):
    return {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("P::")}


def test_dit_forward_fp32_matches_reference(golden_dit):
    g = golden_dit
    cfg, P = _cfg(g), _params(g)
    out = R.dit_forward(cfg, P, torch.from_numpy(g["ids"]), torch.from_numpy(g["modality"]), mode="fp32")
    ref = torch.from_numpy(g["ref_logits_fp32"])
    assert (out - ref).abs().max().item() < 2e-5


def test_dit_forward_bf16_mode_close_to_reference(golden_dit):
    g = golden_dit
    cfg, P = _cfg(g), _params(g)
    out = R.dit_forward(cfg, P, torch.from_numpy(g["ids"]), torch.from_numpy(g["modality"]), mode="bf16").float()
    # bf16 autocast emulation vs the reference's CPU-autocast run and vs its fp32 run
    assert (out - torch.from_numpy(g["ref_logits_cpu_bf16"])).abs().max().item() < 0.06
    assert (out - torch.from_numpy(g["ref_logits_fp32"])).abs().max().item() < 0.06


def test_rope_tables(golden_dit):
    g = golden_dit
    cfg = _cfg(g)
    c, s = R.rope_table_2d(cfg.head_dim, cfg.img_length)
    assert np.array_equal(c.numpy(), g["rope_cos_img"]) and np.array_equal(s.numpy(), g["rope_sin_img"])


def test_sample_t_and_noise(golden_fns):
    g = golden_fns
    t = R.sample_t(torch.from_numpy(g["sample_t_u"]))
    assert np.array_equal(t.numpy(), g["sample_t_ref"])
    s, ds = R.loglinear_noise(t)
    assert np.array_equal(s.numpy(), g["sigma_ref"]) and np.array_equal(ds.numpy(), g["dsigma_ref"])


def test_q_xt_bit_exact(golden_fns, golden_dit):
    g = golden_fns
    mi = int(golden_dit["cfg"][7])
    xt, mv, _ = R.q_xt(torch.from_numpy(g["qxt_x0"]), torch.from_numpy(g["qxt_mc"]), torch.from_numpy(g["qxt_rand"]), mi)
    assert np.array_equal(xt.numpy(), g["qxt_ref"]) and np.array_equal(mv.numpy(), g["qxt_move_ref"])
    mod = torch.from_numpy(g["qxtm_mod"])
    mm = torch.stack([mod == 0, mod == 1], -1)
    xt, mv, ign = R.q_xt(torch.from_numpy(g["qxtm_x0"]), torch.from_numpy(g["qxtm_mc"]), torch.from_numpy(g["qxtm_rand"]), mi,
                         modality_mask=mm, mask_entire_modality=0.9, rand_txt=torch.from_numpy(g["qxtm_rt"]),
                         rand_img=torch.from_numpy(g["qxtm_ri"]))
    assert np.array_equal(xt.numpy(), g["qxtm_ref"]) and np.array_equal(ign.numpy(), g["qxtm_ignore_ref"])


def test_subs_parameterization(golden_fns, golden_dit):
    g = golden_fns
    tv, mi = int(golden_dit["cfg"][6]), int(golden_dit["cfg"][7])
    lg = torch.from_numpy(g["subs_logits"])
    xt, mod = torch.from_numpy(g["subs_xt"]), torch.from_numpy(g["subs_modality"])
    assert np.array_equal(R.subs_parameterization(lg, xt, mod, mi, tv).numpy(), g["subs_f32_ref_xt"])
    assert np.array_equal(R.subs_parameterization(lg, None, mod, mi, tv).numpy(), g["subs_f32_ref_noxt"])
    assert np.array_equal(R.subs_parameterization(lg.bfloat16(), xt, mod, mi, tv).float().numpy(), g["subs_bf16_ref_xt"])


def test_loss_reads_only_masked_rows_and_their_own_vocabulary_block(golden_fns, golden_dit):
    """The claim behind trainer.b200_masked_head / b200_split_head, checked on the oracle that the test above pins bit-exactly to
    the reference's `_subs_parameterization` (model.py:621-658) and loss (model.py:967-1057): the training loss does not depend on
    the logits of unmasked or padded rows, nor on a row's logits outside its modality's vocabulary block -- the loss is bit-identical
    when those entries are replaced by garbage, and its gradient there is exactly zero."""
    g = golden_fns
    tv, mi = int(golden_dit["cfg"][6]), int(golden_dit["cfg"][7])
    lg0 = torch.from_numpy(g["subs_logits"]).float()
    xt, mod = torch.from_numpy(g["subs_xt"]), torch.from_numpy(g["subs_modality"])
    B, N, V = lg0.shape
    gen = torch.Generator().manual_seed(5)
    own = torch.where(mod == 0, torch.randint(0, tv - 1, (B, N), generator=gen), torch.randint(tv, V, (B, N), generator=gen))
    x0 = torch.where(xt == mi, own, xt)
    am = torch.ones(B, N, dtype=torch.bool)
    am[0, N - 3:] = False
    t = torch.linspace(0.2, 0.9, B)
    masked = (xt == mi) & am
    assert masked.any() and (~masked).any()
    cols = torch.arange(V)
    foreign = torch.where((mod == 0)[..., None], cols >= tv, cols < tv)                   # [B,N,V]: the other modality's block
    unread = (~masked)[..., None] | foreign
    for dtype in (torch.float32, torch.bfloat16):
        lg = lg0.to(dtype).clone().requires_grad_(True)
        loss = R.diffusion_loss(R.subs_parameterization(lg, xt, mod, mi, tv), x0, t, mod, am)["loss"]
        loss.backward()
        assert torch.isfinite(loss) and lg.grad[masked].abs().sum() > 0
        assert torch.count_nonzero(lg.grad[unread]) == 0
        junk = (torch.randn(lg0.shape, generator=gen) * 50).to(dtype)
        lg2 = torch.where(unread, junk, lg0.to(dtype))
        loss2 = R.diffusion_loss(R.subs_parameterization(lg2, xt, mod, mi, tv), x0, t, mod, am)["loss"]
        assert torch.equal(loss2, loss.detach()), (dtype, float(loss2), float(loss))


def test_compute_loss(golden_fns, golden_dit):
    g, gd = golden_fns, golden_dit
    cfg, P = _cfg(gd), _params(gd)
    x0, am = torch.from_numpy(g["loss_x0"]), torch.from_numpy(g["loss_am"])
    mod = torch.from_numpy(gd["modality"])
    for tag, w, snr in (("loss_w6_snr0", 0.6, None), ("loss_w5_snr5", 0.5, 5.0)):
        out = R.training_loss(cfg, P, x0, mod, am, torch.from_numpy(g["loss_u_t"]), torch.from_numpy(g["loss_rand_move"]),
                              mode="fp32", img_loss_weight=w, softmin_snr=snr)
        ref = g[tag + "_ref"]
        assert abs(out["loss"].item() - ref[0]) < 1e-4 * max(1, abs(ref[0]))
        assert abs(out["txt_loss"].item() - ref[1]) < 1e-4 * max(1, abs(ref[1]))
        assert np.allclose(out["nlls"].numpy(), g[tag + "_nlls_ref"], rtol=1e-4, atol=1e-3)


def test_samplers_bit_exact(golden_fns, golden_dit):
    g = golden_fns
    mi = int(golden_dit["cfg"][7])
    probs = torch.from_numpy(g["sc_probs"])
    assert np.array_equal(R.sample_categorical(probs, torch.from_numpy(g["sc_u"])).numpy(), g["sc_ref"])
    x, t, dt = torch.from_numpy(g["ddpm_x"]), torch.from_numpy(g["ddpm_t"]), float(g["ddpm_dt"])
    assert np.array_equal(R.ddpm_caching_update(x, t, dt, probs.clone(), torch.from_numpy(g["ddpm_u"]), mi).numpy(), g["ddpm_cache_ref"])
    assert np.array_equal(R.ddpm_update(x, t, dt, probs.clone(), torch.from_numpy(g["ddpm_u3"]), mi).numpy(), g["ddpm_ref"])
    sch = R.adap_sche(x, 8, mi)
    assert np.array_equal(sch.numpy(), g["sche_ref"])
    mg = R.maskgit_update(x, t, probs, torch.from_numpy(g["mg_pred"]), torch.from_numpy(g["mg_gumbel"]), sch[:, 2], mi, r_temp=10)
    assert np.array_equal(mg.numpy(), g["mg_ref"])


def test_samplers_from_logits_bit_exact(golden_sampler_logits):
    """oracle chain SUBS(logits).exp() -> ddpm_cache / ddpm / maskgit against the reference's outputs on its own noise draws
    (the Exp(1) tensor inside torch.multinomial and the np.random.gumbel draw are stored in the fixture)."""
    g = golden_sampler_logits
    V, tv, mi = [int(v) for v in g["cfg"]]
    lg, mod, x = torch.from_numpy(g["logits"]), torch.from_numpy(g["modality"]), torch.from_numpy(g["x"])
    t, dt = torch.from_numpy(g["t"]), float(g["dt"])
    p = R.subs_parameterization(lg, x, mod, mi, tv).exp()
    assert np.array_equal(R.ddpm_caching_update(x, t, dt, p.clone(), torch.from_numpy(g["cache_u"]), mi).numpy(), g["cache_ref"])
    assert np.array_equal(R.ddpm_update(x, t, dt, p.clone(), torch.from_numpy(g["ddpm_u"]), mi).numpy(), g["ddpm_ref"])
    sch = R.adap_sche(x, 8, mi)
    assert np.array_equal(sch.numpy(), g["schedule"])
    for step in (0, 4, 7):
        e, gum = torch.from_numpy(g[f"mg_e_{step}"]), torch.from_numpy(g[f"mg_gumbel_{step}"])
        assert np.array_equal(R.multinomial_from_exponential(p, e).numpy(), g[f"mg_pred_{step}"])
        out = R.maskgit_update_from_noise(x, t, p, e, gum, sch[:, step], mi, r_temp=10)
        assert np.array_equal(out.numpy(), g[f"mg_ref_{step}"])


def test_attention_caching_cycle_matches_reference(golden_attn_cache, golden_dit):
    """oracle restatement of the inference attention-caching cycle (model_eval.py:2297-2367, dit.py:793-812) against the
    unmodified reference DIT's logits and cache contents (step 0 full, step 1 masked + store, step 2 text-only + update)."""
    g, gd = golden_attn_cache, golden_dit
    cfg, P = _cfg(gd), _params(gd)
    txt, N = cfg.txt_length, cfg.length
    x0, x1, x2, mod = (torch.from_numpy(g[k]) for k in ("x0", "x1", "x2", "modality"))
    sl = slice(None, txt)
    cache = {}
    m0 = R.dit_forward(cfg, P, x0, mod, mode="fp32")
    m1 = R.dit_forward(cfg, P, x1, mod, mode="fp32", attn_mask=R.caching_step_mask(txt, N), kv_cache=cache, cache_op="store")
    m2 = R.dit_forward(cfg, P, x2[:, sl], mod[:, sl], mode="fp32", kv_cache=cache, cache_op="update", update_slice=sl)
    for got, key in ((m0, "ref_step0"), (m1, "ref_step1"), (m2, "ref_step2")):
        assert np.abs(got[:, :, ::7].numpy() - g[key]).max() < 3e-5, key
    assert np.abs(cache[1]["k"][:, :, ::5, ::3].numpy() - g["ref_cache_k_blk1"]).max() < 1e-5
    assert np.abs(cache[1]["v"][:, :, ::5, ::3].numpy() - g["ref_cache_v_blk1"]).max() < 1e-5


def test_ddpm_forward_and_cfg(golden_fns, golden_dit):
    g, gd = golden_fns, golden_dit
    cfg, P = _cfg(gd), _params(gd)
    mod = torch.from_numpy(gd["modality"])
    x = torch.from_numpy(g["ddpm_x"])
    lg = R.dit_forward(cfg, P, x, mod, mode="fp32")
    p = R.subs_parameterization(lg, x, mod, cfg.mask_index, cfg.text_vocab_size).exp()
    assert np.allclose(p.numpy(), g["ddpmfwd_p_ref"], rtol=1e-4, atol=1e-6)
    xc, um = torch.from_numpy(g["cfg_x"]), torch.from_numpy(g["cfg_unmask"])
    xu = xc.clone()
    xu[um] = cfg.mask_index
    t = torch.from_numpy(g["ddpm_t"]).squeeze(-1)
    comb = R.cfg_combine(R.dit_forward(cfg, P, xc, mod), R.dit_forward(cfg, P, xu, mod), t, 2.5)
    pc = R.subs_parameterization(comb, None, mod, cfg.mask_index, cfg.text_vocab_size).exp()
    assert np.allclose(pc.numpy(), g["cfg_p_ref"], rtol=2e-4, atol=1e-6)


def test_torch_eager_module_matches_restated_oracle():
    """oracle/torch_eager.py (the library-op restatement timed as "the reference's torch-SDPA path" by bench.py) computes
    the same function as oracle/restated.py (pinned against the unmodified reference above)."""
    from oracle import restated as R
    from oracle import torch_eager as TE
    cfg = R.OracleConfig(128, 4, 2, 16, 16, 64 + 33, 33, 32)
    P = R.init_params(cfg, seed=1)
    m = TE.EagerDIT(cfg, dropout=0.0)
    missing, unexpected = m.load_state_dict(P, strict=True)
    assert not missing and not unexpected
    ids, modality = R.synthetic_batch(3, 16, 16, cfg.text_vocab_size, cfg.vocab_size, seed=5)
    want = R.dit_forward(cfg, P, ids, modality, mode="fp32")
    got = m(ids, modality)
    assert torch.allclose(got, want, rtol=1e-4, atol=1e-5), (got - want).abs().max()
    # training-step restatement runs and back-propagates in eager mode
    m.train()
    loss = TE.reference_style_loss(m, ids, modality, torch.ones_like(ids, dtype=torch.bool), cfg.mask_index, cfg.text_vocab_size,
                                   autocast=False, generator=torch.Generator().manual_seed(0))
    loss.backward()
    assert torch.isfinite(loss) and m.blocks[0].mlp[0].weight.grad.abs().sum() > 0


def _icfg(g):
    import dataclasses
    D, H, L, N, V, tv, mi = [int(v) for v in g["cfg"]]
    return dataclasses.replace(R.OracleConfig(D, H, L, N - 256, 256, V, tv, mi), require_sample_ids=True), N


def test_interleaved_tables_and_forward_match_reference(golden_interleaved):
    """data.require_sample_ids (BASELINE cfg5): per-image-block RoPE tables, img_count_embedding ordinals and the packed-sample
    forward with the document mask vs the reference's add_img_data_to_blocks / add_txt_data_to_blocks / FlexAttention run."""
    g = golden_interleaved
    cfg, N = _icfg(g)
    modality, sid = torch.from_numpy(g["modality"]), torch.from_numpy(g["sample_ids"])
    cos, sin, ordinal = R.interleaved_token_tables(cfg, modality, sid)
    assert np.array_equal(cos.numpy(), g["ref_cos"]) and np.array_equal(sin.numpy(), g["ref_sin"])
    assert np.array_equal(ordinal.numpy(), g["ref_ordinal"])
    assert ordinal.max() >= 1 and (ordinal == -1).any()          # fixture has second images and table-less blocks
    P = _params(g)
    out = R.dit_forward(cfg, P, torch.from_numpy(g["ids"]), modality, mode="fp32", sample_ids=sid)
    ref = torch.from_numpy(g["ref_logits_fp32"])
    valid = sid != -1
    assert (out - ref)[valid].abs().max().item() < 5e-5


def test_interleaved_q_xt_oracle_and_host_logic_bit_exact(golden_interleaved):
    """model.py:483-522 — whole-block masking of packed batches.  (1) the oracle restatement fed the recorded draws and
    (2) the product's torch host function `unidisc_b200.model.q_xt_general` replaying the recorded CPU seed must both
    reproduce the reference's xt / move / ignore masks bit for bit."""
    g = golden_interleaved
    mi = int(g["cfg"][6])
    x0, mc = torch.from_numpy(g["qxt_x0"]), torch.from_numpy(g["qxt_mc"])
    modality, sid = torch.from_numpy(g["modality"]), torch.from_numpy(g["sample_ids"])
    xt, mv, ign = R.q_xt_interleaved(x0, mc, torch.from_numpy(g["qxt_rand"]), mi, modality, sid, 0.2,
                                     torch.from_numpy(g["qxt_rand_blocks"]))
    assert np.array_equal(xt.numpy(), g["qxt_ref"]) and np.array_equal(mv.numpy(), g["qxt_move_ref"])
    assert np.array_equal(ign.numpy(), g["qxt_ignore_ref"])
    from unidisc_b200.model import q_xt_general
    trainer = dict(mask_entire_modality=0.2, interleaved=True)
    torch.manual_seed(int(g["qxt_seed"]))
    xt2, ign2, _, _, mv2 = q_xt_general(x0, mc, mi, trainer, backbone_training=True, training=True,
                                        batch=dict(modality=modality, sample_ids=sid))
    assert np.array_equal(xt2.numpy(), g["qxt_ref"]) and np.array_equal(mv2.numpy(), g["qxt_move_ref"])
    assert np.array_equal(ign2.numpy(), g["qxt_ignore_ref"])


def test_time_conditioning_matches_reference(golden_timecond):
    """config.time_conditioning (adaLN shift / scale / gate from sigma, dit.py:966-1031, 1083-1091) incl. the all-text quirk."""
    import dataclasses
    g = golden_timecond
    cfg = dataclasses.replace(_cfg(g), time_conditioning=True)
    P = _params(g)
    sigma = torch.from_numpy(g["sigma"])
    out = R.dit_forward(cfg, P, torch.from_numpy(g["ids"]), torch.from_numpy(g["modality"]), mode="fp32", sigma=sigma)
    assert (out - torch.from_numpy(g["ref_logits_fp32"])).abs().max().item() < 3e-5
    out_t = R.dit_forward(cfg, P, torch.from_numpy(g["ids_txt"]), torch.from_numpy(g["modality_txt"]), mode="fp32", sigma=sigma[:2])
    assert (out_t - torch.from_numpy(g["ref_logits_txt_fp32"])).abs().max().item() < 3e-5
    # bf16-emulating mode stays close to the fp32 reference
    out_bf = R.dit_forward(cfg, P, torch.from_numpy(g["ids"]), torch.from_numpy(g["modality"]), mode="bf16", sigma=sigma).float()
    assert (out_bf - torch.from_numpy(g["ref_logits_fp32"])).abs().max().item() < 0.08


def _cfg1(g):
    """Parameters of the BASELINE.json configs[0]-sized fixture are rebuilt from their seed (37 M values, not committed)."""
    from oracle.gen_golden import cfg1_params
    ocfg, P = cfg1_params(int(g["param_seed"][0]))
    assert [int(v) for v in g["cfg"]] == [ocfg.hidden_size, ocfg.n_heads, ocfg.n_blocks, ocfg.txt_length, ocfg.img_length,
                                          ocfg.vocab_size, ocfg.text_vocab_size, ocfg.mask_index]
    chk = sum(v.double().sum().item() for v in P.values())
    assert abs(chk - float(g["param_checksum"][0])) < 1e-6 * max(1.0, abs(chk)), "seeded parameter generator changed"
    return ocfg, P


def test_cfg1_reference_parity_run(golden_cfg1):
    """BASELINE.json configs[0]: DiT-S depth 6 / dim 384 / 6 heads, seq_len 256 (60 text + 196 = 14x14 image tokens), batch 4,
    fp32 CPU eager, real vocabulary — the oracle against logits and compute_loss outputs of the unmodified reference."""
    g = golden_cfg1
    ocfg, P = _cfg1(g)
    x0, modality = torch.from_numpy(g["x0"]), torch.from_numpy(g["modality"])
    # q_xt on the stored draws is bit-exact
    t = R.sample_t(torch.from_numpy(g["u_t"]))
    sigma, _ = R.loglinear_noise(t)
    xt, _, _ = R.q_xt(x0, 1 - torch.exp(-sigma[:, None]), torch.from_numpy(g["rand_move"]), ocfg.mask_index)
    assert np.array_equal(xt.numpy(), g["xt"])
    logits = R.dit_forward(ocfg, P, xt, modality, mode="fp32")
    assert (logits[:, ::4, ::61] - torch.from_numpy(g["ref_logits_sub"])).abs().max().item() < 5e-5
    assert (logits.max(dim=-1).values - torch.from_numpy(g["ref_logits_rowmax"])).abs().max().item() < 5e-5
    s, a = g["ref_logits_sum"]
    assert abs(logits.double().abs().sum().item() - a) < 1e-6 * a
    assert abs(logits.double().sum().item() - s) < 1e-6 * a
    agree = (logits.argmax(dim=-1).numpy() == g["ref_logits_argmax"]).mean()
    assert agree > 0.999, agree                      # (near-ties may flip at 1e-6)
    out = R.training_loss(ocfg, P, x0, modality, torch.from_numpy(g["loss_am"]), torch.from_numpy(g["loss_u_t"]),
                          torch.from_numpy(g["loss_rand_move"]), mode="fp32", img_loss_weight=0.6)
    ref = g["loss_ref"]
    assert abs(out["loss"].item() - ref[0]) < 1e-4 * max(1.0, abs(ref[0]))
    assert abs(out["txt_loss"].item() - ref[1]) < 1e-4 and abs(out["img_loss"].item() - ref[2]) < 1e-4
    assert torch.allclose(out["nlls"], torch.from_numpy(g["loss_nlls_ref"]), rtol=1e-4, atol=1e-3)


def test_update_batch_matches_reference():
    """Batch contract: unidisc_b200.model.update_batch against the outputs of the reference's own
    Diffusion.update_batch (model.py:157-395) on tokenised dataloader batches — every tensor identical, dtypes included."""
    import os
    from oracle.gen_golden import _ub_cases, _ub_config
    from oracle.ref_loader import to_attrdict
    from unidisc_b200.model import update_batch
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "update_batch.npz"))
    cases, tv = _ub_cases()
    assert tv == int(g["text_vocab_size"][0])
    for name, (batch, over, txt, img) in cases.items():
        for k, v in batch.items():                                   # the seeded inputs are the ones the fixture was made from
            assert np.array_equal(v.numpy(), g[f"{name}::in::{k}"]), (name, k)
        out = update_batch({k: v.clone() for k, v in batch.items()}, to_attrdict(_ub_config(txt, img, over)), text_vocab_size=tv,
                           device=torch.device("cpu"))
        want = {k.split("::")[2]: g[k] for k in g.files if k.startswith(f"{name}::out::")}
        assert sorted(want) == sorted(k for k, v in out.items() if isinstance(v, torch.Tensor)), name
        for k, w in want.items():
            assert out[k].numpy().dtype == w.dtype and np.array_equal(out[k].numpy(), w), (name, k)
        for k in [k for k in g.files if k.startswith(f"{name}::meta::")]:
            assert np.array_equal(out["interleaved_metadata"][k.split("::")[2]].numpy(), g[k]), (name, k)
    # image ids are shifted into the joint vocabulary, padding is excluded, modality -1 is folded into text
    o = update_batch({k: v.clone() for k, v in cases["pretok_sid"][0].items()},
                     to_attrdict(_ub_config(12, 16, cases["pretok_sid"][1])), text_vocab_size=tv, device=torch.device("cpu"))
    assert int(o["input_ids"][:, 12:].min()) >= tv and int(o["modality"].min()) == 0
    assert bool((o["sample_ids"][~o["attention_mask"]] == -1).all())
    with pytest.raises(NotImplementedError):
        update_batch(dict(img=torch.zeros(1, 3, 8, 8), input_ids=torch.zeros(1, 28), modality=torch.ones(1, 28)),
                     to_attrdict(_ub_config(12, 16, {})), text_vocab_size=tv, device=torch.device("cpu"))


def test_first_hitting_update_matches_reference():
    """First-hitting sampler step (model_eval.py:3004-3043): unidisc_b200.model.first_hitting_select on the reference's draws
    (categorical sample of p_x0, then the random choice of which masked positions to reveal) — bit-exact tokens."""
    import os
    from unidisc_b200.model import first_hitting_select
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "first_hitting.npz"))
    mi = int(g["mask_index"][0])
    x, probs, sched = torch.from_numpy(g["x"]), torch.from_numpy(g["probs"]), torch.from_numpy(g["schedule"])
    for step in (0, 3, 5):
        xs = R.sample_categorical(probs, torch.from_numpy(g[f"u_{step}"]))
        out = first_hitting_select(x.clone(), xs, sched[:, step], torch.from_numpy(g[f"rv_{step}"]), mi)
        assert np.array_equal(out.numpy(), g[f"ref_{step}"]), step
        newly = (out != x)
        assert bool((x[newly] == mi).all())                                   # only masked positions change
        want = torch.minimum(sched[:, step], (x == mi).sum(-1))
        assert torch.equal(newly.sum(-1), want.to(newly.sum(-1).dtype))      # exactly the scheduled number per row
    # nothing to reveal -> unchanged
    assert torch.equal(first_hitting_select(x.clone(), x, torch.zeros(x.shape[0], dtype=torch.int32), torch.rand(x.shape), mi), x)
