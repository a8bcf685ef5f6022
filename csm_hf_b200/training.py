"""Training forward + backward of CSMModel on B200.

Mirrors `CSMModel.forward(input_ids, attention_mask, labels=...)` of the reference (modeling_csm.py:292-482) as it is
used by `CSMTrainer.compute_loss` (train.py:303-326): returns a `CSMOutput` whose `loss` is a scalar tensor that
`loss.backward()` differentiates w.r.t. every parameter of the module -- the forward AND the backward run in
`csm_train_step` (csrc/csm_train.cu: tcgen05 GEMMs, flash attention forward/backward, fused element-wise kernels); torch
autograd only carries the finished parameter gradients to `.grad`.  No PyTorch op is on the compute path and there is
no fallback: without the CUDA library this raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch

from . import native, rope
from .config import CSMConfig, CSMOutput
from .synthetic import state_dict_shapes

_LAYER_KEYS = ["self_attn.q_proj.weight", "self_attn.k_proj.weight", "self_attn.v_proj.weight",
               "self_attn.o_proj.weight", "mlp.gate_proj.weight", "mlp.up_proj.weight", "mlp.down_proj.weight",
               "input_layernorm.weight", "post_attention_layernorm.weight"]


def _weights_struct(cfg: CSMConfig, tensors: Dict[str, torch.Tensor], keep: list) -> native.Weights:
    """The C view (include/csm_b200.h CsmWeights) of a name -> bf16 device tensor mapping in the reference's layout."""
    def ptr(name):
        t = tensors[name]
        if t.dtype != torch.bfloat16 or not t.is_contiguous() or t.device.type != "cuda":
            raise ValueError(f"{name}: training needs contiguous bf16 CUDA tensors (got {t.dtype}, {t.device})")
        return t.data_ptr()

    def layer_ptrs(prefix, n):
        arr = (C.c_void_p * (n * native.W_PER_LAYER))()
        for l in range(n):
            for j, k in enumerate(_LAYER_KEYS):
                arr[l * native.W_PER_LAYER + j] = ptr(f"{prefix}.layers.{l}.{k}")
        keep.append(arr)
        return C.cast(arr, C.POINTER(C.c_void_p))

    w = native.Weights()
    w.text_embeddings = ptr("text_embeddings.weight")
    w.audio_embeddings = ptr("audio_embeddings.weight")
    w.projection = ptr("projection.weight")
    w.codebook0_head = ptr("codebook0_head.weight")
    w.audio_head = ptr("audio_head")
    w.backbone_norm = ptr("backbone.norm.weight")
    w.decoder_norm = ptr("decoder.norm.weight")
    w.backbone_layers = layer_ptrs("backbone", cfg.backbone_config.num_hidden_layers)
    w.decoder_layers = layer_ptrs("decoder", cfg.decoder_config.num_hidden_layers)
    return w


class TrainEngine:
    """Owns one CsmTrain (workspace for up to max_tokens = B*S tokens and max_frames decoder frames per step)."""

    def __init__(self, cfg: CSMConfig, device: torch.device, max_tokens: int, max_frames: int, max_seq: int):
        if device.type != "cuda":
            raise RuntimeError("CSM training runs only on a CUDA (B200, sm_100a) device")
        self.lib = native.load()
        self.cfg, self.device = cfg, device
        self.max_tokens, self.max_frames, self.max_seq = max_tokens, max_frames, max_seq
        keep = []

        def llama_shape(d, n_pos):
            cos, sin = rope.tables(d.head_dim, d.rope_theta, d.rope_scaling, n_pos)
            keep.extend([cos, sin])
            s = native.LlamaShape()
            s.hidden, s.inter, s.layers = d.hidden_size, d.intermediate_size, d.num_hidden_layers
            s.heads, s.kv_heads, s.eps = d.num_attention_heads, d.num_key_value_heads, d.rms_norm_eps
            s.rope_cos, s.rope_sin, s.n_pos = cos.data_ptr(), sin.data_ptr(), n_pos
            return s

        sh = native.Shapes()
        sh.text_vocab, sh.audio_vocab, sh.n_codebooks = cfg.text_vocab_size, cfg.audio_vocab_size, cfg.audio_num_codebooks
        sh.backbone = llama_shape(cfg.backbone_config, max_seq)
        sh.decoder = llama_shape(cfg.decoder_config, 33)
        self.ctx = C.c_void_p()
        with torch.cuda.device(device):
            rc = self.lib.csm_train_create(C.byref(sh), max_tokens, max_frames, C.byref(self.ctx))
            try:
                self._check(rc)
            except Exception:
                self.close()
                raise

    def _check(self, rc):
        if rc >= 0:
            return rc
        msg = self.lib.csm_train_last_error(self.ctx).decode() if self.ctx else "csm training error"
        if rc in (native.CSM_EINVAL, native.CSM_ECAPACITY):
            raise ValueError(msg)
        if rc == native.CSM_EUNSUPPORTED:
            raise NotImplementedError(msg)
        raise RuntimeError(msg)

    def close(self):
        if getattr(self, "ctx", None) is not None and self.ctx:
            self.lib.csm_train_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def launches(self) -> int:
        return int(self.lib.csm_train_launches(self.ctx))

    def step(self, params: Dict[str, torch.Tensor], grads: Optional[Dict[str, torch.Tensor]], ids: torch.Tensor,
             mask: Optional[torch.Tensor], labels: torch.Tensor):
        """-> (losses [3] host floats, n_frames, last_h [B,Hb], c0_logits [B,V]); `grads` tensors are overwritten."""
        B, S = ids.shape[:2]
        keep: list = []
        w = _weights_struct(self.cfg, params, keep)
        g = _weights_struct(self.cfg, grads, keep) if grads is not None else None
        losses = (C.c_float * 3)()
        nfr = C.c_int(0)
        last_h = torch.empty(B, self.cfg.backbone_config.hidden_size, dtype=torch.bfloat16, device=self.device)
        c0 = torch.empty(B, self.cfg.audio_vocab_size, dtype=torch.bfloat16, device=self.device)
        with torch.cuda.device(self.device):
            st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            rc = self.lib.csm_train_step(self.ctx, C.byref(w), C.byref(g) if g is not None else None, ids.data_ptr(),
                                         mask.data_ptr() if mask is not None else None, labels.data_ptr(), B, S, losses,
                                         C.byref(nfr), last_h.data_ptr(), c0.data_ptr(), st)
        self._check(rc)
        return [float(x) for x in losses], int(nfr.value), last_h, c0

    def step_begin(self, params, grads, ids, mask, labels, split_layer: int):
        """First call of a split step (csm_train_step_begin); -> (n_frames, last_h, c0_logits).  step_end() must follow."""
        B, S = ids.shape[:2]
        keep: list = []
        w = _weights_struct(self.cfg, params, keep)
        g = _weights_struct(self.cfg, grads, keep)
        nfr = C.c_int(0)
        last_h = torch.empty(B, self.cfg.backbone_config.hidden_size, dtype=torch.bfloat16, device=self.device)
        c0 = torch.empty(B, self.cfg.audio_vocab_size, dtype=torch.bfloat16, device=self.device)
        self._pending_keep = (keep, w, g, ids, mask, labels)      # alive until step_end returns
        with torch.cuda.device(self.device):
            st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            rc = self.lib.csm_train_step_begin(self.ctx, C.byref(w), C.byref(g), ids.data_ptr(),
                                               mask.data_ptr() if mask is not None else None, labels.data_ptr(), B, S,
                                               split_layer, C.byref(nfr), last_h.data_ptr(), c0.data_ptr(), st)
        self._check(rc)
        return int(nfr.value), last_h, c0

    def step_end(self):
        """Second call of a split step: -> losses [3] (synchronises the stream)."""
        losses = (C.c_float * 3)()
        with torch.cuda.device(self.device):
            st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            rc = self.lib.csm_train_step_end(self.ctx, losses, st)
        self._pending_keep = None
        self._check(rc)
        return [float(x) for x in losses]

    def debug(self, name: str, dtype=torch.bfloat16) -> torch.Tensor:
        """Host copy of a named intermediate of the last step (tests)."""
        n = C.c_longlong(0)
        self._check(self.lib.csm_train_debug(self.ctx, name.encode(), None, 0, C.byref(n)))
        out = torch.empty(n.value // torch.empty((), dtype=dtype).element_size(), dtype=dtype)
        if n.value:
            self._check(self.lib.csm_train_debug(self.ctx, name.encode(), out.data_ptr(), n.value, C.byref(n)))
        return out


class _CSMLoss(torch.autograd.Function):
    """loss = f(parameters): forward and backward both ran inside csm_train_step; backward hands the gradients over."""

    @staticmethod
    def forward(ctx, loss_value: torch.Tensor, grads: List[torch.Tensor], *params):
        ctx.grads = grads
        return loss_value.clone()

    @staticmethod
    def backward(ctx, gout):
        grads = ctx.grads
        ctx.grads = None
        if grads is None:
            raise RuntimeError("the CSM training loss can be differentiated once (its gradients were computed with the forward)")
        scale = float(gout)
        if scale != 1.0:
            for g in grads:
                if g is not None:
                    g.mul_(scale)
        return (None, None) + tuple(grads)


def training_forward(model, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor], labels: torch.Tensor,
                     return_dict: Optional[bool] = True):
    """CSMModel.forward with labels (modeling_csm.py:292-482): -> CSMOutput(loss, backbone_loss, decoder_loss,
    last_hidden_state, logits).  Gradients are produced when autograd is enabled and a parameter requires grad."""
    cfg = model.config
    dev = model.device
    if input_ids.dim() != 3 or input_ids.shape[-1] != cfg.audio_num_codebooks + 1:
        raise ValueError(f"input_ids must be [B,S,{cfg.audio_num_codebooks + 1}]")
    if labels.shape != input_ids.shape:
        raise ValueError(f"labels {tuple(labels.shape)} must match input_ids {tuple(input_ids.shape)}")
    B, S = input_ids.shape[:2]
    ids = input_ids.to(device=dev, dtype=torch.int64).contiguous()
    lab = labels.to(device=dev, dtype=torch.int64).contiguous()
    mask = None
    if attention_mask is not None:
        if attention_mask.shape != input_ids.shape:
            raise ValueError("attention_mask must match input_ids")
        mask = (attention_mask != 0).to(device=dev, dtype=torch.int32).contiguous()   # (the processor pads with float32 masks)
    nq, V = cfg.audio_num_codebooks, cfg.audio_vocab_size
    # ids index the embedding tables and labels index the logits on the device: out-of-range values must not get there
    if int(ids[..., :nq].min()) < 0 or int(ids[..., :nq].max()) >= V or int(ids[..., nq].min()) < 0 \
            or int(ids[..., nq].max()) >= cfg.text_vocab_size:
        raise IndexError("input_ids out of range of the embedding tables")
    lab_a = lab[..., :nq]
    if bool(((lab_a < 0) & (lab_a != -100)).any()) or int(lab_a.max()) >= V:
        raise IndexError(f"labels must be -100 or in [0, {V})")
    names = list(state_dict_shapes(cfg).keys())
    params = dict(model.named_parameters())
    missing = [k for k in names if k not in params]
    if missing:
        raise RuntimeError(f"parameters not loaded: {missing[:4]}")
    plist = [params[k] for k in names]
    want = torch.is_grad_enabled() and any(p.requires_grad for p in plist)
    n_frames_cap = int((lab[:, :, : cfg.audio_num_codebooks] != -100).all(dim=2).sum().item()) if B * S else 0
    eng: Optional[TrainEngine] = getattr(model, "_train_engine", None)
    if (eng is None or eng.device != dev or eng.max_tokens < B * S or eng.max_frames < max(n_frames_cap, 1)
            or eng.max_seq < S):
        if eng is not None:
            eng.close()
        eng = TrainEngine(cfg, dev, B * S, max(n_frames_cap, 1), S)
        model._train_engine = eng
    grads = None
    ddp = getattr(model, "_ddp", None) if want else None
    if want:
        # one flat bf16 buffer, the gradient tensors are views into it: data-parallel training all-reduces the buffer
        # without packing / unpacking copies.  Layout: first the gradients that become final LAST in the backward
        # (embedding tables, backbone layers below the split), then the rest -- the two halves are averaged separately,
        # the second one while the backward of the first is still running (enable_data_parallel below).
        split = ddp["split"] if ddp else 0
        late = [k for k in names if k.endswith("embeddings.weight") or
                (k.startswith("backbone.layers.") and int(k.split(".")[2]) < split)]
        order = late + [k for k in names if k not in set(late)]
        pmap = dict(zip(names, plist))
        sizes = {k: -(-pmap[k].numel() // 8) * 8 for k in order}                     # 16-byte aligned views
        flat = torch.empty(sum(sizes.values()), dtype=torch.bfloat16, device=dev)
        grads, off, late_end = {}, 0, 0
        for k in order:
            grads[k] = flat[off:off + pmap[k].numel()].view(pmap[k].shape)
            off += sizes[k]
            if k in late:
                late_end = off
        model._grad_flat = flat
    pdata = {k: p.data for k, p in zip(names, plist)}
    if ddp:
        import torch.distributed as dist
        group, side = ddp["group"], ddp["stream"]
        cur = torch.cuda.current_stream(dev)
        op = dist.ReduceOp.AVG if dist.get_backend(group) == "nccl" else dist.ReduceOp.SUM
        n_frames, last_h, c0 = eng.step_begin(pdata, grads, ids, mask, lab, split)
        ev = torch.cuda.Event()
        ev.record(cur)
        flat.record_stream(side)
        with torch.cuda.stream(side):           # averaged on the side stream while the rest of the backward runs
            side.wait_event(ev)
            dist.all_reduce(flat[late_end:], op=op, group=group)
        losses = eng.step_end()
        with torch.cuda.stream(side):
            if late_end:
                dist.all_reduce(flat[:late_end], op=op, group=group)
            if op == dist.ReduceOp.SUM:
                flat.div_(dist.get_world_size(group))
        cur.wait_stream(side)
        model._grads_reduced = True
    else:
        losses, n_frames, last_h, c0 = eng.step(pdata, grads, ids, mask, lab)
    vals = torch.tensor(losses, dtype=torch.float32, device=dev)
    loss = vals[0]
    if want:
        glist = [grads[k] if p.requires_grad else None for k, p in zip(names, plist)]
        loss = _CSMLoss.apply(vals[0], glist, *plist)
        model._drop_engine()   # (the generation engine's packed weights go stale once the optimizer steps)
    out = CSMOutput(last_hidden_state=last_h, logits=c0, loss=loss, backbone_loss=vals[1],
                    decoder_loss=vals[2].to(torch.bfloat16))   # (bf16 like the reference's: CE on bf16 logits, :464-467)
    if return_dict is False:
        return (loss, last_h, c0)
    return out


def enable_data_parallel(model, group=None, split_layer: Optional[int] = None):
    """Data-parallel training with the gradient all-reduce overlapped with the backward: every later
    `model(..., labels=...)` step averages its gradients over `group` itself (two NCCL all-reduces on a side stream, the
    first one -- decoder, heads and the backbone layers >= split_layer -- while the backbone layers below split_layer are
    still in their backward), and `dist.allreduce_gradients(model)` becomes a no-op for that step.
    split_layer defaults to a quarter of the backbone's layers."""
    L = model.config.backbone_config.num_hidden_layers
    split = max(0, min(L, L // 4 if split_layer is None else split_layer))
    model._ddp = {"group": group, "split": split, "stream": torch.cuda.Stream(device=model.device)}
    return model
