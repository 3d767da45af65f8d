"""GPU parity of `config.time_conditioning` (adaLN shift / scale / gate from sigma; reference dit.py:415-449, 966-1031,
1083-1091, 229-304): forward against the golden logits of the unmodified reference (tests/golden/timecond.npz) and the
bf16-mode oracle, training step (loss + every gradient incl. adaLN_modulation / sigma_map) against autograd through the
oracle.  Tolerances as in test_model_gpu.py."""
import dataclasses

import pytest
import torch

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16


def dev():
    return torch.device("cuda", 0)


def _model(g):
    from oracle import restated as R
    from unidisc_b200.config import make_config
    from unidisc_b200.dit import DIT
    D, H, L, txt, img, V, tv, mi = [int(v) for v in g["cfg"]]
    cfg = make_config("small", hidden_size=D, n_blocks=L, n_heads=H, txt_length=txt, img_length=img, image_vocab_size=V - tv,
                      text_vocab_size=tv, time_conditioning=True)
    m = DIT(cfg, vocab_size=V, text_vocab_size=tv, mask_index=mi).to(dev())
    P = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("P::")}
    r = m.load_state_dict(P)
    assert not r.missing_keys and not r.unexpected_keys
    return m, P, dataclasses.replace(R.OracleConfig(D, H, L, txt, img, V, tv, mi), time_conditioning=True)


def test_timecond_forward_matches_reference_golden(golden_timecond):
    from oracle import restated as R
    g = golden_timecond
    m, P, ocfg = _model(g)
    m.eval()
    sigma = torch.from_numpy(g["sigma"])
    for ik, mk, rk, sg in (("ids", "modality", "ref_logits_fp32", sigma), ("ids_txt", "modality_txt", "ref_logits_txt_fp32", sigma[:2])):
        ids, mod = torch.from_numpy(g[ik]), torch.from_numpy(g[mk])
        with torch.no_grad():
            out = m(ids.to(dev()), sg.to(dev()), modality=mod.to(dev())).float().cpu()
        ref32 = torch.from_numpy(g[rk])
        orc = R.dit_forward(ocfg, P, ids, mod, mode="bf16", sigma=sg).float()
        e_orc, e_ref = (out - orc).abs(), (out - ref32).abs()
        print(f"[{ik}] max|cuda-oracle_bf16|={e_orc.max():.4f} mean={e_orc.mean():.5f}; max|cuda-reference_fp32|={e_ref.max():.4f}")
        assert e_orc.max() < 4e-2 and e_orc.mean() < 4e-3
        assert e_ref.max() < 8e-2
    with pytest.raises(ValueError):
        m(ids.to(dev()), None, modality=mod.to(dev()))


@pytest.mark.parametrize("dropout", [0.0, 0.1])
def test_timecond_training_step_vs_oracle(dropout):
    from oracle import restated as R
    from unidisc_b200 import ops
    from unidisc_b200.config import make_config
    from unidisc_b200.model import Diffusion
    D, H, L, txt, img, tv, iv = 256, 4, 2, 64, 64, 257, 255
    cfg = make_config("small", hidden_size=D, n_blocks=L, n_heads=H, txt_length=txt, img_length=img, image_vocab_size=iv,
                      text_vocab_size=tv, img_loss_weight=0.6, dropout=dropout, time_conditioning=True)
    torch.manual_seed(0)
    model = Diffusion(cfg, device=dev())
    model.train()
    net = model.backbone
    gen = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for n, p in net.named_parameters():
            if "adaLN_modulation" in n:                      # zero-initialised in the reference: make them matter
                p.copy_(((torch.rand(p.shape, generator=gen) - 0.5) * (1.2 if n.endswith("weight") else 0.5)).to(dev()))
    net.mark_weights_updated()
    V, mi = model.vocab_size, model.mask_index
    B, N = 4, txt + img
    ids, modality = R.synthetic_batch(B, txt, img, tv, V, seed=3)
    am = torch.ones(B, N, dtype=torch.bool)
    batch = dict(input_ids=ids.to(dev()), modality=modality.to(dev()), attention_mask=am.to(dev()))
    torch.manual_seed(11)
    out = model.compute_loss(batch)
    out.loss.backward()
    torch.cuda.synchronize()
    ks = None
    if dropout > 0:
        base = net._dropout_calls * L
        ks = [ops.dropout_scales(B * N, D, dropout, net.dropout_seed, base + i, dev()).view(B, N, D).cpu() for i in range(L)]
    torch.manual_seed(11)
    u_t = torch.rand(B, device=dev()).cpu()
    rand_move = torch.rand(B, N, device=dev()).cpu()
    P = {k: v.detach().float().cpu().clone().requires_grad_(True) for k, v in net.state_dict().items()}
    assert any("adaLN_modulation" in k for k in P) and "sigma_map.mlp.0.weight" in P
    ocfg = dataclasses.replace(R.OracleConfig(D, H, L, txt, img, V, tv, mi), time_conditioning=True)
    ref_bf = R.training_loss(ocfg, {k: v.detach() for k, v in P.items()}, ids, modality, am, u_t, rand_move, mode="bf16", drop_scales=ks)
    plain = R.training_loss(R.OracleConfig(D, H, L, txt, img, V, tv, mi), {k: v.detach() for k, v in P.items()}, ids, modality, am, u_t,
                            rand_move, mode="bf16", drop_scales=ks)
    ref32 = R.training_loss(ocfg, P, ids, modality, am, u_t, rand_move, mode="fp32", drop_scales=ks)
    ref32["loss"].backward()
    got = float(out.loss.detach())
    print(f"loss cuda={got:.6f} oracle_bf16={float(ref_bf['loss']):.6f} oracle_fp32={float(ref32['loss']):.6f} (unconditioned {float(plain['loss']):.6f})")
    assert abs(got - float(ref_bf["loss"])) < 1e-3 * max(1.0, abs(float(ref_bf["loss"]))) + 2e-3
    assert abs(float(ref_bf["loss"]) - float(plain["loss"])) > 5e-4          # the conditioning really changes the result
    worst = ("", 0.0)
    for name, p in net.named_parameters():
        gref = P[name].grad
        assert gref is not None, name
        gg = p.grad.detach().float().cpu()
        den = gref.norm().item()
        rel = (gg - gref).norm().item() / max(den, 1e-8)
        if rel > worst[1] and den >= 1e-6:
            worst = (name, rel)
        lim = 1e-1 if ("q_norm" in name or "k_norm" in name) else 3e-2
        if "sigma_map" in name:
            lim = 6e-2                                       # four bf16 roundings between sigma and c
        assert rel < lim or den < 1e-6, f"grad {name}: rel L2 err {rel:.4f} (|ref|={den:.3e})"
    print("worst grad rel err:", worst)
