#!/usr/bin/env python
"""bench.py -- audio frames/s of CSM greedy generation (BASELINE.json metric).

A "step" is one CSMModel.generate() of `--frames` new frames after a `--ctx`-frame context at
`--batch` sequences per GPU (default: BASELINE.json configs[1] = csm-1b bf16, 2048-frame
context, 200 new frames, batch 1, one B200).  Prints ONE JSON line (rank 0).

  value     whole-job frames/s (all ranks), inputs resident in HBM, prefill included
  e2e       the same call through the host-buffer C-ABI entry (csm_generate_host): pinned H2D of
            ids+mask and D2H of the frames inside the timed region
  roofline  the persistent decode-frame kernel (csm_stream_kernel): algorithmic bytes per
            launch / CUDA-event time per launch, against MEASURED_PEAKS.json
  config.points   the other batch sizes of the metric ("batch 1/8/32"), measured in the same
            invocation with the same engine code: frames/s, decode ms/frame, roofline fraction,
            prefill time against the tensor roofline.  Under torchrun with N ranks the batch-8
            point is BASELINE.json configs[3] scaled to N GPUs (8 sequences per GPU, N x 8 total).
  verified  after the timed region the SAME engine is checked, teacher-forced, against the
            committed fixture minted from the reference at this configuration
            (tests/golden/csm1b_t2048_b1_fp32.pt): logits within the stated tolerance, greedy ids
            identical wherever the reference's margin decides them
  cpu_baseline / --impl reference: the reference's own CPU path (modeling_csm.py staged under
            oracle/_ref by oracle/stage_ref.py; the oracle port if that copy is absent), fp32, all
            host threads, on a bounded sample (prefill + a few decode frames) PROJECTED to the
            200-frame workload -- the JSON says so
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# parameter counts of csm-1b
P_BB, P_C0, P_DEC, P_AH, P_PROJ = 973_078_528, 4_200_448, 111_149_056, 65_106_944, 2_097_152


def algorithmic_bytes(B, T):
    """bf16 bytes one decode step must move: every weight once (decoder 31x: its passes are
    sequentially dependent) + the cached K/V of every sequence + the gathered embedding rows."""
    return 2 * (P_BB + P_C0 + 31 * P_DEC + P_AH + P_PROJ) + B * T * 32768 + B * 65 * 4096


def prefill_flops(B, T):
    """2*B*T*P_bb + causal attention + last-position head."""
    return 2.0 * B * T * P_BB + B * 16 * 4 * 32 * 64 * T * T / 2.0 + 2.0 * B * P_C0


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            j = json.load(f)
        return (float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)",
                float(j.get("bf16_tflops_sustained", 1400.0)), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)")
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)", 1400.0, "fallback (B200_PROFILING.md ~1.4 PFLOP/s sustained)"


def ncu_traffic(batch):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE frame launch at this batch size (profiles/traffic.json: taken
    from the committed ncu --set full captures), or None when that batch size was not captured."""
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.isfile(tp):
        return None
    with open(tp) as f:
        return json.load(f).get(f"b{batch}")


def training_point(cfg, sd, dev, rank, world, seq, tpeak, note, steps=5):
    """ms per training step (CSMModel.forward(labels=...) + backward, csrc/csm_train.cu) and its tensor-roofline fraction."""
    import torch
    import torch.distributed as dist
    from csm_hf_b200.modeling import CSMModel
    from csm_hf_b200.synthetic import make_training_batch
    ok, err, ms, out, F = 1, None, 0.0, None, 0
    try:
        tm = CSMModel(cfg, sd, device=dev)
        tm.requires_grad_(True)
        ids, mask, labels = [t.to(dev) for t in make_training_batch(cfg, 1, seq, seed=100 + rank, text_frames=16)]
        F = int((labels[:, :, :32] != -100).all(dim=2).sum())
        for _ in range(2):                                                # warm-up steps, local
            tm.zero_grad(set_to_none=True)
            out = tm(input_ids=ids, attention_mask=mask, labels=labels)
            out.loss.backward()
        torch.cuda.synchronize()
    except Exception as e:   # noqa: BLE001
        ok, err = 0, f"{type(e).__name__}: {e}"
    if world > 1:
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = int(flag)
    if not ok:
        return {"error": err or "another rank failed"}
    from csm_hf_b200.dist import allreduce_gradients
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        # (only now, after every rank has completed local steps: from here on each step contains collectives)
        from csm_hf_b200.training import enable_data_parallel
        enable_data_parallel(tm)         # gradient all-reduce on a side stream, overlapped with the backward
        for _ in range(2):
            tm.zero_grad(set_to_none=True)
            out = tm(input_ids=ids, attention_mask=mask, labels=labels)
            out.loss.backward()
            allreduce_gradients(tm)
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        tm.zero_grad(set_to_none=True)
        out = tm(input_ids=ids, attention_mask=mask, labels=labels)
        out.loss.backward()
        if world > 1:
            allreduce_gradients(tm)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from train_bench import flops_forward
    fl = 3 * flops_forward(cfg, 1, seq, F)
    res = {"workload": f"csm-1b fwd+bwd, seq_len {seq}, one sequence per GPU x {world} GPU(s), 1/16 decoder amortisation "
                       f"({F} frames), gradients all-reduced over NCCL" if world > 1 else
                       f"csm-1b fwd+bwd, seq_len {seq}, one sequence, 1/16 decoder amortisation ({F} frames)",
           "ms_per_step": ms, "steps": steps, "tokens_per_s": world * seq / (ms / 1000.0), "loss": float(out.loss.detach()),
           "roofline": {"bound": "tensor", "achieved": fl / (ms / 1000.0) / 1e12, "peak": tpeak, "unit": "TFLOP/s",
                        "frac": fl / (ms / 1000.0) / 1e12 / tpeak, "algorithmic_flops_per_step_per_gpu": fl},
           "gpu_launches_per_step": tm._train_engine.launches() // (steps + 2)}
    note(f"training point: {ms:.1f} ms/step")
    del tm
    torch.cuda.empty_cache()
    return res


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def next_row(toks):
    import torch
    B = toks.shape[0]
    ids = torch.cat([toks.to(torch.long), torch.zeros(B, 1, dtype=torch.long)], dim=1).unsqueeze(1)
    m = torch.zeros(B, 1, 33, dtype=torch.int32)
    m[:, :, :32] = 1
    return ids, m


class CpuReference:
    """The reference's CPU path for the bounded sample: the real modeling_csm.py (oracle/_ref or /root/reference)
    behind the harness shims of oracle/ref_harness.py, else the oracle port."""

    def __init__(self, cfg, sd_fp32):
        import torch
        from oracle import ref_harness as R   # bench.py executes oracle/ only here: as the timed CPU baseline
        torch.set_num_threads(os.cpu_count() or 1)
        self.kind = "port"
        self.model = None
        if R.reference_available() and not os.environ.get("BENCH_CPU_PORT"):
            try:
                self.model = R.build_reference_model(cfg, sd_fp32, torch.float32)
                self.kind = "reference"
            except Exception as exc:  # noqa: BLE001 -- e.g. a transformers version the reference cannot import under
                sys.stderr.write(f"[bench] reference import failed ({type(exc).__name__}: {exc}); timing the oracle port\n")
        if self.model is None:
            from oracle.csm_oracle import CSMOracle
            self.oracle = CSMOracle(cfg, sd_fp32, torch.float32)

    def sample(self, ids, mask, frames_sample):
        """-> (seconds of prefill + first frame, seconds per decode frame)"""
        import torch
        import warnings
        with torch.inference_mode(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if self.model is not None:
                t0 = time.perf_counter()
                out = self.model.generate_frame(ids, mask, temperature=1.0, topk=1, past_key_values=None, use_cache=True,
                                                return_dict=True)
                t_first = time.perf_counter() - t0
                kv, toks = out.past_key_values, out.samples
                t1 = time.perf_counter()
                for _ in range(frames_sample):
                    rid, rm = next_row(toks)
                    out = self.model.generate_frame(rid, rm, temperature=1.0, topk=1, past_key_values=kv, use_cache=True,
                                                    return_dict=True)
                    kv, toks = out.past_key_values, out.samples
                return t_first, (time.perf_counter() - t1) / frames_sample
            cache = self.oracle.new_cache(1, ids.shape[1] + frames_sample + 1)
            t0 = time.perf_counter()
            toks, _, _ = self.oracle.generate_frame(ids, mask, cache)
            t_first = time.perf_counter() - t0
            t1 = time.perf_counter()
            for _ in range(frames_sample):
                rid, rm = next_row(toks)
                toks, _, _ = self.oracle.generate_frame(rid, rm, cache)
            return t_first, (time.perf_counter() - t1) / frames_sample


def note(msg):
    """Progress on stderr (BENCH_VERBOSE=1): where a run is, if it ever stops."""
    if os.environ.get("BENCH_VERBOSE"):
        sys.stderr.write(f"[bench {time.strftime('%H:%M:%S')}] {msg}\n")
        sys.stderr.flush()


def reference_arm(a, cfg, workload):
    import signal
    import torch
    from csm_hf_b200.synthetic import make_context, make_state_dict

    def _too_slow(signum, frame):   # a host that cannot finish the bounded CPU sample in 15 minutes
        print(json.dumps({"impl": "reference", "unavailable": "CPU reference sample did not finish within 900 s"}))
        sys.stdout.flush()
        os._exit(0)

    signal.signal(signal.SIGALRM, _too_slow)
    signal.alarm(900)
    t_start = time.perf_counter()
    ref = CpuReference(cfg, make_state_dict(cfg, seed=0))
    ids, mask = make_context(cfg, 1, a.ctx, seed=1234)
    vals = []
    n_warm = 1                                                    # one warm-up pass is enough on the CPU
    budget_s = float(os.environ.get("BENCH_REF_BUDGET_S", "420"))  # the whole arm ends within a few minutes
    for i in range(n_warm + a.steps):
        t_first, t_frame = ref.sample(ids, mask, a.cpu_sample_frames)
        if i >= n_warm:
            vals.append((a.frames / (t_first + (a.frames - 1) * t_frame), t_first, t_frame))
        if vals and time.perf_counter() - t_start > budget_s:
            break
    v = statistics.mean(x[0] for x in vals)
    src = "modeling_csm.py (the reference itself, staged copy)" if ref.kind == "reference" else "oracle port of modeling_csm.py"
    sample = (f"{src}, fp32, torch CPU: {a.ctx}-frame prefill+frame ({statistics.mean(x[1] for x in vals):.2f} s) + "
              f"{a.cpu_sample_frames} decode frames ({statistics.mean(x[2] for x in vals):.3f} s each), PROJECTED to "
              f"{a.frames} frames, batch 1; {len(vals)} samples")
    print(json.dumps({
        "impl": "reference", "metric": "audio_frames_per_s", "value": v, "unit": "frames/s", "n_gpus": a.gpus,
        "steps": len(vals), "warmup": n_warm, "ms_per_step": 1000.0 * a.frames / v, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "projected": True,
        "config": {"workload": workload.replace(f"batch={a.batch} per GPU", "batch=1"),
                   "note": "ms_per_step is the projected time of one 200-frame generate(): each timed sample runs the "
                           "prefill and a few decode frames of it"},
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": ref.kind,
                         "sample": sample, "projected": True},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def verify_against_fixture(model, dev):
    """Teacher-forced check of the engine that was just timed against the reference-minted fixture of this
    configuration (csm-1b, 2048-frame context, batch 1): -> dict for the JSON line."""
    import torch
    path = os.path.join(ROOT, "tests", "golden", "csm1b_t2048_b1_fp32.pt")
    if not os.path.isfile(path):
        return {"ok": False, "why": "fixture missing"}
    from csm_hf_b200.synthetic import make_context
    g = torch.load(path, weights_only=False)
    r = g["recipe"]
    ids, mask = make_context(model.config, r["batch"], r["ctx_frames"], seed=r["ctx_seed"])
    rel, worst, ok = 0.05, 0.0, True
    kv, rid, rm = None, ids, mask
    for f in range(g["frames"].shape[1]):
        out = model.generate_frame(rid, rm, temperature=0, past_key_values=kv, force_tokens=g["frames"][:, f],
                                   return_codebook_logits=True)
        kv = out.past_key_values
        for got, want in ((out.logits.cpu(), g["c0_logits"][f]), (out.codebook_logits.cpu(), g["cb_logits"][f])):
            want = want.float()
            scale = float(want.abs().max())
            err = float((got.float() - want).abs().max()) / scale
            worst = max(worst, err)
            top2 = torch.topk(want, 2, dim=-1).values
            decided = (top2[..., 0] - top2[..., 1]) > 2 * rel * scale
            ok = ok and err <= rel and not bool(((got.float().argmax(-1) != want.argmax(-1)) & decided).any())
        rid, rm = next_row(g["frames"][:, f])
    return {"ok": bool(ok), "fixture": "tests/golden/csm1b_t2048_b1_fp32.pt (reference fp32 CPU run)",
            "frames": int(g["frames"].shape[1]), "max_err_of_logit_range": worst, "tolerance": rel,
            "how": "teacher-forced with the reference's ids after the timed region, same engine"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1, help="sequences per GPU of the headline value")
    ap.add_argument("--ctx", type=int, default=2048)
    ap.add_argument("--frames", type=int, default=200)
    ap.add_argument("--points", default="1,8,32", help="batch sizes per GPU reported under config.points ('' = none)")
    ap.add_argument("--point-steps", type=int, default=3)
    ap.add_argument("--no-train-point", action="store_true",
                    help="skip the training-step point (BASELINE config #5: fwd+bwd, seq_len 4096, one sequence per GPU)")
    ap.add_argument("--train-seq", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--cpu-sample-frames", type=int, default=4)
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    workload = (f"csm-1b random-init, greedy (temperature=0), {a.ctx}-frame synthetic context, {a.frames} new frames, "
                f"batch={a.batch} per GPU, KV cache")

    import torch
    from csm_hf_b200.config import CSMConfig
    from csm_hf_b200.synthetic import make_context, make_state_dict
    cfg = CSMConfig()

    if a.impl == "reference":
        if rank == 0:
            reference_arm(a, cfg, workload)
        return

    # ------------------------------------------------------------------ our arm
    import torch.distributed as dist
    from csm_hf_b200.dist import all_gather_frames, generate_sharded
    from csm_hf_b200.modeling import CSMModel
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    want_cpu = (world == 1 and not a.no_cpu_baseline)
    cpu_baseline = None
    if want_cpu:
        # the CPU baseline is the reference arm run in a child process BEFORE any GPU work, with a time limit: it
        # cannot disturb (or be disturbed by) the GPU measurement, and a slow host cannot stall the bench
        note("cpu baseline (child process) starts")
        cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "1", "--ctx",
               str(a.ctx), "--frames", str(a.frames), "--cpu-sample-frames", str(a.cpu_sample_frames)]
        try:
            res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
            ref = json.loads(res.stdout.strip().splitlines()[-1])
            cpu_baseline = ref["cpu_baseline"]
        except Exception as exc:   # noqa: BLE001 -- a baseline that cannot be measured is reported, not fatal
            cpu_baseline = {"value": None, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
                            "sample": f"not measured: {type(exc).__name__}"}
        note("cpu baseline done")
    sd = {k: v.to(dev) for k, v in make_state_dict(cfg, seed=0, dtype=torch.bfloat16).items()}
    peak, peak_src, tpeak, tpeak_src = measured_peaks()
    t_mean = a.ctx + (a.frames + 1) / 2.0                     # mean cached length over the decode frames

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(batch, steps, warmup, headline=False):
        """Device-resident inputs, `steps` timed generate() calls after `warmup`: -> (model, dict)."""
        model = CSMModel(cfg, sd, device=dev, max_batch=batch, max_ctx=a.ctx + a.frames + 8)
        GB = batch * world
        ids, mask = make_context(cfg, GB, a.ctx, seed=1234)
        d_ids, d_mask = ids.to(dev), mask.to(dev)
        eng = model.engine(batch, a.ctx + a.frames)

        def step():
            if world > 1:
                return generate_sharded(model, d_ids, d_mask, max_new_frames=a.frames, temperature=0, stop_on_all_zeros=False)
            return model.generate(d_ids, d_mask, max_new_frames=a.frames, temperature=0, stop_on_all_zeros=False)

        for i in range(warmup):
            out = step()
            torch.cuda.synchronize()
        assert tuple(out.shape) == (GB, a.frames, 32)
        launches0 = eng.info(4)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dec_ms, dec_n = 0.0, 0
        fence()
        ev0.record()
        marks = []
        for _ in range(steps):
            out = step()
            ms, n = model.last_decode_ms()
            dec_ms += ms
            dec_n += n
            marks.append(torch.cuda.Event(enable_timing=True))
            marks[-1].record()
        ev1.record()
        fence()
        per = sorted(b.elapsed_time(c) for b, c in zip([ev0] + marks[:-1], marks))
        t = torch.tensor([ev0.elapsed_time(ev1), dec_ms / max(dec_n, 1), per[len(per) // 2] if len(per) % 2 else
                          0.5 * (per[len(per) // 2 - 1] + per[len(per) // 2])], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, dec_per, ms_median = [float(x) for x in t.cpu()]
        ms_step = ms_total / steps if headline else ms_median   # (points: the median step, robust to a one-off stall)
        abytes = algorithmic_bytes(batch, t_mean)
        achieved = abytes / (dec_per / 1000.0) / 1e9 if dec_per > 0 else 0.0
        pre_ms = max(ms_step - dec_per * (a.frames - 1), 1e-6)       # prefill + the first frame's launch
        ptf = prefill_flops(batch, a.ctx) / (pre_ms / 1000.0) / 1e12
        res = {"batch_per_gpu": batch, "global_batch": GB, "value": GB * a.frames / (ms_step / 1000.0), "unit": "frames/s",
               "ms_per_step": ms_step, "ms_per_step_median": ms_median, "steps": steps, "decode_ms_per_frame": dec_per,
               "decode_frames_per_s_per_gpu": batch * 1000.0 / dec_per if dec_per > 0 else 0.0,
               "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                            "algorithmic_bytes_per_launch": abytes, "traffic": ncu_traffic(batch)},
               "prefill": {"ms_incl_first_frame": pre_ms, "bound": "tensor", "achieved": ptf, "peak": tpeak,
                           "unit": "TFLOP/s", "frac": ptf / tpeak},
               "launches": int(eng.info(4) - launches0), "out0": out[0, :3].cpu() if rank == 0 else None}
        return model, res, (ids, mask)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    note("headline point starts")
    model, head, (ids, mask) = measure(a.batch, a.steps, max(a.warmup, 3), headline=True)
    note("timed device steps done")
    # end to end through host buffers (same process, same engine)
    lo = rank * a.batch
    h_ids, h_mask = ids[lo:lo + a.batch].contiguous().pin_memory(), mask[lo:lo + a.batch].contiguous().pin_memory()
    GB = a.batch * world
    for _ in range(2):
        model.generate(h_ids, h_mask, max_new_frames=a.frames, temperature=0, stop_on_all_zeros=False)
    fence()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        fr = model.generate(h_ids, h_mask, max_new_frames=a.frames, temperature=0, stop_on_all_zeros=False)
        if world > 1:
            all_gather_frames(fr.to(dev), GB)
    fence()
    note("timed host-buffer steps done")
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([e2e_s * 1000.0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.cpu()[0])
    verified = None
    if rank == 0 and not a.no_verify and a.ctx == 2048:
        vm = model if a.batch == 1 else None
        if vm is not None:
            verified = verify_against_fixture(vm, dev)
            note(f"verified: {verified}")
    # the other batch sizes of the metric, same invocation
    points = []
    want_points = [int(x) for x in a.points.split(",") if x.strip()]
    for b in want_points:
        if b == a.batch:
            pt = dict(head)
        else:
            model._drop_engine()
            note(f"point batch {b} starts")
            pm, pt, _ = measure(b, a.point_steps, 3)
            if rank == 0 and verified is None and not a.no_verify and a.ctx == 2048 and b == 1:
                verified = verify_against_fixture(pm, dev)
            pm._drop_engine()
            del pm
        pt.pop("out0", None)
        points.append(pt)
    # BASELINE config #5: one training step = forward with labels + backward (+ NCCL gradient all-reduce
    # when N > 1), csm-1b, seq_len 4096, one sequence per GPU, 1/16 decoder amortisation.  Reported under config.training;
    # never allowed to break the generation line.
    training = None
    if not a.no_train_point and os.environ.get("BENCH_NO_TRAIN") != "1":
        model._drop_engine()
        training = training_point(cfg, sd, dev, rank, world, a.train_seq, tpeak, note)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    e2e_value = GB * a.frames / (e2e_ms / 1000.0 / a.steps)
    traffic = ncu_traffic(a.batch)
    rl = dict(head["roofline"])
    rl.update({"kernel": "csm_stream_kernel (one launch = one frame for the whole batch)", "peak_source": peak_src,
               "traffic": traffic})
    line = {
        "metric": "audio_frames_per_s", "value": head["value"], "unit": "frames/s", "n_gpus": world, "steps": a.steps,
        "warmup": max(a.warmup, 3), "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload, "batch_per_gpu": a.batch, "global_batch": GB, "ctx_frames": a.ctx,
                   "new_frames": a.frames, "parallelism": f"batch-sharded x{world}, weights replicated",
                   "l2": "inputs larger than L2: 3.1 GB of weights are streamed every frame",
                   "decode_ms_per_frame": head["decode_ms_per_frame"],
                   "decode_frames_per_s_per_gpu": head["decode_frames_per_s_per_gpu"],
                   "prefill": dict(head["prefill"], peak_source=tpeak_src),
                   "points": points, "training": training},
        "roofline": rl,
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int(h_ids.numel() * 8 + h_mask.numel() * 4),
                "d2h_bytes_per_step": int(a.batch * a.frames * 32 * 8)},
        "gpu_launches": head["launches"],
        "clocks": clocks,
        "verified": verified,
    }
    if want_cpu:
        line["cpu_baseline"] = cpu_baseline
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
