"""GPU debugging aid (not a test): one training step of the tiny config on the CUDA engine next to the CPU oracle's
autograd, intermediate by intermediate, forward then backward, then every parameter gradient against the
reference-minted fixture.  Usage: python tests/train_debug.py [bf16|fp32-oracle]"""
import os
import sys

os.environ.setdefault("CSM_TRAIN_SNAP", "1")

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from csm_hf_b200.config import tiny_config  # noqa: E402
from csm_hf_b200.modeling import CSMModel  # noqa: E402
from csm_hf_b200.synthetic import make_state_dict, make_training_batch  # noqa: E402
from helpers import dense_grad  # noqa: E402
from oracle.csm_oracle import CSMOracle  # noqa: E402
from oracle.csm_train_oracle import training_forward  # noqa: E402


def rel(a, b):
    a, b = a.float().flatten(), b.float().flatten()
    if a.numel() != b.numel():
        return f"SIZE {a.numel()} vs {b.numel()}"
    scale = b.abs().max().item() + 1e-20
    err = (a - b).abs().max().item() / scale
    cos = torch.nn.functional.cosine_similarity(a, b, dim=0).item() if a.norm() > 0 and b.norm() > 0 else float("nan")
    flag = "" if err < 0.05 else "   <<<<<<"
    return f"max err {err:9.5f} of scale {scale:10.4g}  cos {cos:8.5f}{flag}"


def main():
    fx = torch.load(os.path.join(ROOT, "tests", "golden", "tiny_train_bf16.pt"), weights_only=False)
    r = fx["recipe"]
    cfg = tiny_config()
    sd = make_state_dict(cfg, seed=r["weight_seed"], norm_jitter=r["norm_jitter"])
    ids, mask, labels = make_training_batch(cfg, r["batch"], r["frames"], seed=r["seed"], text_frames=r["text_frames"],
                                            amortization_ratio=r["amortization_ratio"], pad=r["pad"])
    # oracle with every intermediate kept
    osd = {k: v.to(torch.bfloat16).clone().requires_grad_(True) for k, v in sd.items()}
    keep = {}
    loss, bl, dl = training_forward(CSMOracle(cfg, osd, torch.bfloat16), ids, mask, labels, keep)
    loss.backward()
    print(f"oracle    loss {loss.item():.5f} = {bl.item():.5f} + {dl.item():.5f}")
    print(f"reference loss {fx['loss'].item():.5f} = {fx['backbone_loss'].item():.5f} + {fx['decoder_loss'].item():.5f}")

    dev = torch.device("cuda", 0)
    model = CSMModel(cfg, sd, device=dev, max_batch=1, max_ctx=64)
    model.requires_grad_(True)
    out = model(input_ids=ids, attention_mask=mask, labels=labels)
    print(f"engine    loss {out.loss.item():.5f} = {out.backbone_loss.item():.5f} + {float(out.decoder_loss):.5f}")
    eng = model._train_engine
    print("launches", eng.launches())
    names = {"bb.hf": "hf", "bb.h_out": "bb.h_out", "dec.hf": "hdf"}
    order = []
    for tag, L in (("bb", cfg.backbone_config.num_hidden_layers), ("dec", cfg.decoder_config.num_hidden_layers)):
        for l in range(L):
            for n in ("h_in", "hn1", "qkv", "attn", "h_mid", "hn2", "act"):
                order.append((f"{tag}.{l}.{n}", f"{tag}.{l}.{n}"))
        order.append((f"{tag}.h_out", f"{tag}.h_out") if tag == "bb" else ("dec.hf", "hdf"))
        if tag == "bb":
            order.append(("bb.hf", "hf"))
    print("---- forward intermediates (engine vs oracle)")
    for mine, theirs in order:
        if theirs not in keep:
            continue
        print(f"{mine:14s} {rel(eng.debug(mine), keep[theirs].detach())}")
    print("---- backward intermediates")
    border = []
    for tag, L in (("dec", cfg.decoder_config.num_hidden_layers), ("bb", cfg.backbone_config.num_hidden_layers)):
        border.append(("d.bb.hf", "hf") if tag == "bb" else None)
        for l in reversed(range(L)):
            for n in ("act", "hn2", "h_mid", "attn", "hn1"):
                border.append((f"d.{tag}.{l}.{n}", f"{tag}.{l}.{n}"))
        border.append(("d.dec_x0", "dec_in") if tag == "dec" else ("d.bb.x0", "bb.0.h_in"))
    for it in border:
        if it is None:
            continue
        mine, theirs = it
        g = keep[theirs].grad
        if g is None:
            print(f"{mine:14s} (oracle kept no gradient)")
            continue
        print(f"{mine:14s} {rel(eng.debug(mine), g)}")
    out.loss.backward()
    print("---- parameter gradients (engine vs reference fixture)")
    worst = 0.0
    for k, p in model.named_parameters():
        want = dense_grad(fx["grads"][k])
        line = rel(p.grad.cpu(), want)
        print(f"{k:55s} {line}")
    print("done")


if __name__ == "__main__":
    main()
