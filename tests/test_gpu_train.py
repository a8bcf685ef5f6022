"""Training step on the B200: CSMModel.forward(labels=...) + loss.backward() through the
C ABI (csm_train_step) against fixtures minted from the unmodified reference (oracle/make_golden.py --train: the loss
triple and all 43 parameter gradients of the tiny configuration, reference bf16 and fp32) and against the CPU oracle's
autograd (oracle/csm_train_oracle.py) on other batches.

Tolerances (bf16 tensors, fp32 accumulation; the engine shares the reference's rounding points in the forward but not
its accumulation order, and sums embedding / norm-weight gradients in fp32 where autograd adds bf16 partial results):
  * loss, backbone_loss, decoder_loss: 2e-3 relative to the reference's bf16 run (decoder_loss is a bf16 number:
    one bf16 ulp at 4.3 is 0.03), 1e-2 relative to its fp32 run;
  * every parameter gradient: max |err| <= 4 % of the tensor's largest entry and cosine >= 0.9995 against the
    reference's bf16 gradients -- the reference's own bf16-vs-fp32 gradient distance on this batch is 2-3 %
    (tests/test_oracle_golden.py::test_training_oracle_vs_reference measures the CPU restatement at 2.2 %).
"""
import pytest
import torch

from helpers import dense_grad

pytestmark = pytest.mark.gpu

import os  # noqa: E402

from csm_hf_b200.config import tiny_config  # noqa: E402
from csm_hf_b200.synthetic import make_state_dict, make_training_batch  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda", 0)


def fixture_batch(tag):
    fx = torch.load(os.path.join(GOLD, f"tiny_train_{tag}.pt"), weights_only=False)
    r = fx["recipe"]
    cfg = tiny_config()
    sd = make_state_dict(cfg, seed=r["weight_seed"], norm_jitter=r["norm_jitter"])
    ids, mask, labels = make_training_batch(cfg, r["batch"], r["frames"], seed=r["seed"], text_frames=r["text_frames"],
                                            amortization_ratio=r["amortization_ratio"], pad=r["pad"])
    return fx, cfg, sd, ids, mask, labels


def grad_close(got, want, what, rel=0.04, cos_min=0.9995):
    got, want = got.float().flatten().cpu(), want.float().flatten()
    scale = float(want.abs().max())
    if scale == 0.0:
        assert float(got.abs().max()) == 0.0, f"{what}: reference gradient is zero, engine's is not"
        return
    err = float((got - want).abs().max())
    cos = float(torch.nn.functional.cosine_similarity(got, want, dim=0))
    assert err <= rel * scale and cos >= cos_min, f"{what}: max err {err:.3g} of {scale:.3g}, cosine {cos:.6f}"


def test_loss_and_gradients_vs_reference_fixture(dev):
    """modeling_csm.py:367-471 + autograd: loss triple and every parameter gradient of a left-padded, amortised batch."""
    from csm_hf_b200.modeling import CSMModel
    fx, cfg, sd, ids, mask, labels = fixture_batch("bf16")
    model = CSMModel(cfg, sd, device=dev)
    model.requires_grad_(True)
    out = model(input_ids=ids, attention_mask=mask, labels=labels)
    for k in ("loss", "backbone_loss", "decoder_loss"):
        got, want = float(getattr(out, k).detach()), float(fx[k])
        assert abs(got - want) <= 2e-3 * abs(want) + (0.04 if k != "backbone_loss" else 0.0), (k, got, want)
    f32 = torch.load(os.path.join(GOLD, "tiny_train_fp32.pt"), weights_only=False)
    assert abs(float(out.loss) - float(f32["loss"])) <= 1e-2 * float(f32["loss"])
    assert out.loss.dtype == torch.float32 and out.loss.requires_grad
    assert tuple(out.last_hidden_state.shape) == (ids.shape[0], cfg.backbone_config.hidden_size)
    assert tuple(out.logits.shape) == (ids.shape[0], cfg.audio_vocab_size)
    out.loss.backward()
    params = dict(model.named_parameters())
    assert set(params) == set(fx["grads"])
    for k, p in params.items():
        assert p.grad is not None and p.grad.dtype == torch.bfloat16 and p.grad.shape == p.shape, k
        grad_close(p.grad, dense_grad(fx["grads"][k]), k)
    with pytest.raises(RuntimeError):
        out.loss.backward()                       # the gradients were handed over once


@pytest.mark.parametrize("case", ["no_mask", "no_decoder_frames", "first_frame_selected", "longer"])
def test_loss_and_gradients_vs_oracle_autograd(dev, case):
    """Other batch shapes against torch autograd of the CPU restatement (bf16): attention_mask=None (all 33 slots summed,
    modeling_csm.py:328-332), a batch where no frame carries decoder labels (decoder_loss = 0, :468-469), a selected
    frame at t = 0 (h[b, t-1] wraps to the last position, :405-407), and a longer sequence crossing the 64-key blocks."""
    from csm_hf_b200.modeling import CSMModel
    from oracle.csm_train_oracle import loss_and_grads
    cfg = tiny_config()
    sd = make_state_dict(cfg, seed=11, norm_jitter=0.1)
    B, S = (2, 150) if case == "longer" else (3, 20)
    ids, mask, labels = make_training_batch(cfg, B, S, seed=99, text_frames=0 if case == "first_frame_selected" else 3,
                                            amortization_ratio=3 if case != "longer" else 16, pad=0 if case != "longer" else 70)
    if case == "no_mask":
        mask = None
    if case == "no_decoder_frames":
        labels[:, :, 1:32] = -100
    if case == "first_frame_selected":
        labels[1, 0, :32] = ids[1, 0, :32]
    (l, bl, dl), grads = loss_and_grads(cfg, sd, torch.bfloat16, ids, mask, labels)
    model = CSMModel(cfg, sd, device=dev)
    model.requires_grad_(True)
    out = model(input_ids=ids, attention_mask=mask, labels=labels)
    assert abs(float(out.backbone_loss) - float(bl)) <= 2e-3 * float(bl)
    assert abs(float(out.decoder_loss) - float(dl)) <= 2e-3 * float(dl) + 0.04
    if case == "no_decoder_frames":
        assert float(out.decoder_loss) == 0.0
    out.loss.backward()
    for k, p in model.named_parameters():
        grad_close(p.grad, grads[k], f"{case}: {k}")


def test_loss_only_and_determinism(dev):
    """torch.no_grad(): the loss alone (no gradient tensors are touched); two steps on the same batch give the same
    loss bits and the same matrix gradients (only the fp32 atomics of the table / norm gradients may reorder)."""
    from csm_hf_b200.modeling import CSMModel
    fx, cfg, sd, ids, mask, labels = fixture_batch("bf16")
    model = CSMModel(cfg, sd, device=dev)
    with torch.no_grad():
        o0 = model(input_ids=ids, attention_mask=mask, labels=labels)
    assert not o0.loss.requires_grad
    model.requires_grad_(True)
    o1 = model(input_ids=ids, attention_mask=mask, labels=labels)
    o1.loss.backward()
    g1 = {k: p.grad.clone() for k, p in model.named_parameters()}
    model.zero_grad(set_to_none=True)
    o2 = model(input_ids=ids, attention_mask=mask, labels=labels)
    o2.loss.backward()
    assert float(o0.loss) == float(o1.loss) == float(o2.loss)
    for k, p in model.named_parameters():
        if k.endswith("proj.weight") and "layers" in k and "q_proj" not in k:
            assert torch.equal(p.grad, g1[k]), k            # (q_proj: dQ is reduced with fp32 atomics)
        else:
            grad_close(p.grad, g1[k].cpu(), k, rel=0.02, cos_min=0.9999)


def test_sgd_steps_reduce_the_loss_and_generation_sees_new_weights(dev):
    """A few plain SGD steps on one batch through torch.optim (the reference trains through HF Trainer / AdamW,
    train.py:329-508): the loss falls, and generate() afterwards runs on the updated parameters."""
    from csm_hf_b200.modeling import CSMModel
    from csm_hf_b200.synthetic import make_context
    fx, cfg, sd, ids, mask, labels = fixture_batch("bf16")
    model = CSMModel(cfg, sd, device=dev, max_batch=2, max_ctx=64)
    ctx_ids, ctx_mask = make_context(cfg, 1, 6, seed=3)
    before = model.generate(ctx_ids, ctx_mask, max_new_frames=2, temperature=0, stop_on_all_zeros=False)
    model.requires_grad_(True)
    opt = torch.optim.SGD(model.parameters(), lr=0.05)
    losses = []
    for _ in range(6):
        opt.zero_grad(set_to_none=True)
        out = model(input_ids=ids, attention_mask=mask, labels=labels)
        out.loss.backward()
        opt.step()
        losses.append(float(out.loss))
    assert losses[-1] < losses[0] - 0.2, losses
    after = model.generate(ctx_ids, ctx_mask, max_new_frames=2, temperature=0, stop_on_all_zeros=False)
    assert after.shape == before.shape


def test_training_errors_are_loud(dev):
    from csm_hf_b200.modeling import CSMModel
    fx, cfg, sd, ids, mask, labels = fixture_batch("bf16")
    model = CSMModel(cfg, sd, device=dev)
    with pytest.raises(ValueError):
        model(input_ids=ids, attention_mask=mask, labels=labels[:, :-1])
    with pytest.raises(ValueError):
        model(input_ids=ids[:, :, :5], attention_mask=mask, labels=labels)
    bad = ids.clone()
    bad[0, 3, 2] = cfg.audio_vocab_size
    with pytest.raises(IndexError):
        model(input_ids=bad, attention_mask=mask, labels=labels)
    badl = labels.clone()
    badl[1, 5, 0] = cfg.audio_vocab_size + 3
    with pytest.raises(IndexError):
        model(input_ids=ids, attention_mask=mask, labels=badl)


def test_csm1b_loss_and_gradient_norms_vs_reference_fixture(dev):
    """csm-1b dimensions (16 + 4 layers, head dims 64 and 128, vocabulary 2051 padded to 2112 columns, 2 x 96 frames
    left-padded, 1/8 amortisation) against the reference's fp32 CPU forward(labels=...) + backward
    (tests/golden/csm1b_train_fp32.pt, oracle/make_golden.py --train-1b): loss triple within 1 %, the norm of every one
    of the 187 parameter gradients within 6 %, and the stored gradient rows within 8 % of their largest entry
    (bf16 pipeline against an fp32 reference)."""
    from csm_hf_b200.config import CSMConfig
    from csm_hf_b200.modeling import CSMModel
    fx = torch.load(os.path.join(GOLD, "csm1b_train_fp32.pt"), weights_only=False)
    r = fx["recipe"]
    cfg = CSMConfig()
    sd = make_state_dict(cfg, seed=r["weight_seed"], norm_jitter=r["norm_jitter"])
    ids, mask, labels = make_training_batch(cfg, r["batch"], r["frames"], seed=r["seed"], text_frames=r["text_frames"],
                                            amortization_ratio=r["amortization_ratio"], pad=r["pad"])
    model = CSMModel(cfg, sd, device=dev)
    model.requires_grad_(True)
    out = model(input_ids=ids, attention_mask=mask, labels=labels)
    for k in ("loss", "backbone_loss", "decoder_loss"):
        got, want = float(getattr(out, k).detach()), float(fx[k])
        assert abs(got - want) <= 1e-2 * want, (k, got, want)
    out.loss.backward()
    worst = 0.0
    for k, p in model.named_parameters():
        want = fx["grad_norms"][k]
        got = float(p.grad.float().norm())
        worst = max(worst, abs(got - want) / want)
        assert abs(got - want) <= 0.06 * want, (k, got, want)
    for k, rows in fx["grad_samples"].items():
        g = dict(model.named_parameters())[k].grad
        got = g[:, :2, :] if k == "audio_head" else g[:4]
        grad_close(got, rows, f"rows of {k}", rel=0.08, cos_min=0.998)
    del model
    torch.cuda.empty_cache()


def test_split_step_equals_single_call(dev):
    """csm_train_step_begin(split_layer) + csm_train_step_end (the data-parallel overlap path) produce the same losses and
    gradients as the single call, for every split position."""
    from csm_hf_b200.modeling import CSMModel
    from csm_hf_b200.synthetic import state_dict_shapes
    fx, cfg, sd, ids, mask, labels = fixture_batch("bf16")
    model = CSMModel(cfg, sd, device=dev)
    model.requires_grad_(True)
    out = model(input_ids=ids, attention_mask=mask, labels=labels)
    out.loss.backward()
    ref = {k: p.grad.clone() for k, p in model.named_parameters()}
    eng = model._train_engine
    names = list(state_dict_shapes(cfg).keys())
    params = {k: p.data for k, p in model.named_parameters()}
    d_ids, d_mask, d_lab = ids.to(dev), mask.to(dev).to(torch.int32), labels.to(dev)
    for split in range(cfg.backbone_config.num_hidden_layers + 1):
        grads = {k: torch.zeros_like(params[k]) for k in names}
        eng.step_begin(params, grads, d_ids, d_mask, d_lab, split)
        losses = eng.step_end()
        assert abs(losses[0] - float(out.loss.detach())) < 1e-6
        for k in names:
            if k.endswith("proj.weight") and "q_proj" not in k and "layers" in k:
                assert torch.equal(grads[k], ref[k]), (split, k)
            else:
                grad_close(grads[k], ref[k].cpu(), f"split {split}: {k}", rel=0.02, cos_min=0.9999)


# ------------------------------------------------------------------ the tcgen05 attention kernels on their own
def _flash_lib():
    import ctypes as C
    from csm_hf_b200 import native
    lib = native.load()
    vp = C.c_void_p
    lib.csm_flash_tc_launch.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, vp, vp, vp, vp]
    lib.csm_flash_tc_launch.restype = C.c_int
    lib.csm_flash_tc_bwd_launch.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, vp, vp, vp, vp]
    lib.csm_flash_tc_bwd_launch.restype = C.c_int
    return lib


@pytest.mark.parametrize("S,nseq,heads,kv,padded", [(128, 1, 4, 2, False), (200, 2, 4, 1, False), (333, 2, 8, 2, True),
                                                     (1000, 1, 32, 8, False)])
def test_tcgen05_flash_forward_backward_vs_fp32_sdpa(dev, S, nseq, heads, kv, padded):
    """csm_flash_tc_kernel / csm_flash_tc_bwd_kernel (head dim 64) against torch's SDPA and its autograd in fp32 on the same
    bf16 inputs: ragged lengths (S not a multiple of the 128-row tiles), several sequences, GQA ratios 2 / 4 / 8, and a
    left-padded sequence (keys of padded frames hidden; hf sdpa_attention_forward semantics).  Tolerances: output 1 % of
    its scale (P is rounded to bf16 before P V, as in every flash kernel), log-sum-exp 1e-3, gradients 2 % of each
    tensor's largest entry."""
    import ctypes as C
    lib = _flash_lib()
    HD = 64
    W = (heads + 2 * kv) * HD
    g = torch.Generator().manual_seed(S + heads)
    qkv = (torch.randn(nseq * S, W, generator=g) * 0.7).to(torch.bfloat16).to(dev)
    d_out = (torch.randn(nseq * S, heads * HD, generator=g) * 0.5).to(torch.bfloat16).to(dev)
    valid = None
    if padded:
        valid = torch.ones(nseq, S, dtype=torch.uint8)
        valid[0, :37] = 0
        valid = valid.to(dev).contiguous()
    out = torch.zeros(nseq * S, heads * HD, dtype=torch.bfloat16, device=dev)
    lse = torch.zeros(nseq * S, heads, dtype=torch.float32, device=dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    scale = HD ** -0.5
    vptr = valid.data_ptr() if valid is not None else None
    assert lib.csm_flash_tc_launch(qkv.data_ptr(), S, nseq, heads, kv, C.c_float(scale), vptr, out.data_ptr(), lse.data_ptr(), st) == 0
    # fp32 reference with autograd
    x = qkv.float().view(nseq, S, W)
    q = x[..., : heads * HD].reshape(nseq, S, heads, HD).transpose(1, 2).clone().requires_grad_(True)
    k = x[..., heads * HD: (heads + kv) * HD].reshape(nseq, S, kv, HD).transpose(1, 2).clone().requires_grad_(True)
    v = x[..., (heads + kv) * HD:].reshape(nseq, S, kv, HD).transpose(1, 2).clone().requires_grad_(True)
    rep = heads // kv
    s = (q @ k.repeat_interleave(rep, dim=1).transpose(-1, -2)) * scale
    hide = torch.ones(S, S, dtype=torch.bool, device=dev).triu(1)[None, None].expand(nseq, 1, S, S)
    if valid is not None:
        hide = hide | (valid == 0)[:, None, None, :]
    s = s.masked_fill(hide, float("-inf"))
    p = torch.softmax(s, dim=-1)
    p = torch.where(hide.all(dim=-1, keepdim=True), torch.zeros_like(p), p)       # rows that see nothing -> 0
    ref = (p @ v.repeat_interleave(rep, dim=1)).transpose(1, 2).reshape(nseq * S, heads * HD)
    live = ~hide.all(dim=-1).expand(nseq, heads, S).transpose(1, 2).reshape(nseq * S, heads)
    assert float((out.float() - ref).abs().max()) <= 0.01 * float(ref.abs().max())
    ref_lse = torch.logsumexp(s, dim=-1).transpose(1, 2).reshape(nseq * S, heads)
    assert float((lse - ref_lse)[live].abs().max()) <= 1e-3
    ref.backward(d_out.float())
    # backward through the kernel: D = rowsum(dO * O) as the training step computes it
    delta = (d_out.float() * out.float()).view(nseq * S, heads, HD).sum(-1).contiguous()
    dqkv = torch.zeros_like(qkv)
    dq_acc = torch.zeros(nseq * S, heads * HD, dtype=torch.float32, device=dev)
    assert lib.csm_flash_tc_bwd_launch(qkv.data_ptr(), d_out.data_ptr(), lse.data_ptr(), delta.data_ptr(), S, nseq, heads, kv,
                                       C.c_float(scale), vptr, dqkv.data_ptr(), dq_acc.data_ptr(), st) == 0
    torch.cuda.synchronize()
    want_dq = q.grad.transpose(1, 2).reshape(nseq * S, heads * HD)
    want_dk = k.grad.transpose(1, 2).reshape(nseq * S, kv * HD)
    want_dv = v.grad.transpose(1, 2).reshape(nseq * S, kv * HD)
    got_dk = dqkv[:, heads * HD: (heads + kv) * HD].float()
    got_dv = dqkv[:, (heads + kv) * HD:].float()
    for name, got, want in (("dQ", dq_acc, want_dq), ("dK", got_dk, want_dk), ("dV", got_dv, want_dv)):
        err = float((got - want).abs().max())
        assert err <= 0.02 * float(want.abs().max()), f"{name}: max err {err:.4g} of {float(want.abs().max()):.4g}"


def test_fallback_knobs_keep_working(dev):
    """The documented knobs that switch the training step back to its earlier formulations -- mma.sync attention
    (CSM_FLASH_MMA=1) and transposed copies instead of MN-major GEMM operands (CSM_TRAIN_TRANSPOSE=1) -- still pass the
    reference-fixture test (they are read when the library's contexts are created: a fresh process)."""
    import subprocess
    import sys
    env = dict(os.environ, CSM_FLASH_MMA="1", CSM_TRAIN_TRANSPOSE="1")
    here = os.path.abspath(__file__)
    r = subprocess.run([sys.executable, "-m", "pytest", here, "-q", "-m", "gpu", "-k",
                        "loss_and_gradients_vs_reference_fixture or left_padded or first_frame_selected"],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
