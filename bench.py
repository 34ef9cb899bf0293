#!/usr/bin/env python
"""bench.py -- decoder clips/s of the class-query decoder hot path (BASELINE.json metric, configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--mode train|infer] [--impl ours|reference]

A step (default --mode train = BASELINE.json configs[1] "full decoder fwd+bwd bf16, batch 32") = one
TransformerDecoder.forward (cqvad_decoder_train_forward) + its backward (cqvad_decoder_backward: gradients of every decoder
parameter, memory, tgt and refpoints_unsigmoid) over one batch of synthetic clips (AVA22_ViT-B shapes: nq 15, S 14x14, K 80,
6 layers, F 2048), bf16 tensor-core path, loss = sum(w_hs*hs) + sum(w_cls*cls_hs) + sum(w_refs*refs) (SURVEY.md section 8d
Config 2), nn.Dropout(0.1) ON at the nine residual-branch / FFN-hidden sites of every layer pair (Philox masks regenerated in the
backward; the reference arms run the reference modules in eval mode: dropout is extra work only on this arm).  --mode infer times the inference forward + DETR heads (cqvad_decoder_forward); the default
run reports it under "inference_forward"; at N=1 the line also carries "encoder_layer" (one
deformable encoder layer, forward and forward + backward, on the ViT-B/224 pyramid).
  value : whole-job clips/s with inputs already resident in HBM (CUDA events, max over ranks)
  e2e   : same metric through the public API (DecoderEngine.forward_train/backward) with pinned-HOST inputs: H2D copy of the
          step's inputs and D2H read of the step's result (refs + last-layer hs) inside the timed region; the copy of step i+1
          is issued on a copy stream while step i computes (double-buffered loader; every step still copies its own inputs)
  roofline     : the dominant kernel (tcgen05 implicit-GEMM 3x3 conv: forward and data-gradient launches), timed live with
                 CUDA events on the launch stream (library profiler scopes), algorithmic FLOPs / duration vs the measured peak
  cpu_baseline : the unmodified reference decoder (baseline/_ref) on the host cores, torch-CPU all threads, bounded sample
                 (rank 0, N=1 only); kind "port" (oracle restatement) only if that install is missing
`--impl reference` times the UNMODIFIED reference decoder on the host cores (byte-for-byte install under the git-ignored
baseline/_ref, oracle/install_ref.py -- it travels to the GPU box; /root/reference is never read here), exactly --warmup +
--steps steps of one clip each.  "reference_gpu_eager" (N=1): the same unmodified reference run eagerly on the B200 itself
(fp32 with TF32 off = the parity oracle's arithmetic, and bf16 autocast) -- BASELINE.md section 5.1.
Multi-GPU: one process per GPU (torchrun), clips sharded across ranks (weak scaling: --batch clips per GPU); train mode
all-reduces the decoder gradients once per step (NCCL), infer mode all-gathers the per-clip detections.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

DROPOUT_P = 0.1                        # --dropout
CFG = "ava_vitb"                       # --config: BASELINE.json configs[1] by default; configs[2..3] shapes selectable
# decoder forward GFLOP per clip (reference flop count); fwd + bwd = 3x
FWD_GFLOP = {"ava_vitb": 148.18, "ava_csn152": 187.98, "ucf_vitb": 2147.7, "jhmdb_vitb": 1162.8}
DEFAULT_BATCH = {"ava_vitb": 32, "ava_csn152": 32, "ucf_vitb": 1, "jhmdb_vitb": 1}   # clips/GPU (UCF / JHMDB: 32 / 40 frames per clip)
CFG_NAME = {"ava_vitb": "AVA22_ViT-B", "ava_csn152": "AVA22_CSN152", "ucf_vitb": "UCF_ViT-B", "jhmdb_vitb": "JHMDB_ViT-B"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))


def _host_case():
    from oracle import synth
    cfg = synth.CONFIGS[CFG]
    W = synth.make_decoder_weights(cfg["K"], cfg["layers"], cfg["F"], seed=0)
    inp = synth.make_decoder_inputs(CFG, 1, seed=0)
    lw = synth.make_loss_weights(cfg, 1, seed=0)
    return cfg, W, inp, lw


def host_step_fn(mode):
    """One clip of the workload on the host cores, all threads.  Returns (step, kind, description):
    kind "reference" = the UNMODIFIED reference decoder (+ restated head lines) imported from the git-ignored install
    baseline/_ref (oracle/install_ref.py); only when that install is absent, kind "port" = the oracle
    restatement (torch-CPU for the training step, numpy for inference)."""
    from oracle import ref_runner
    if ref_runner.available():
        n = ref_runner.cpu_threads()
        step = ref_runner.step_fn(CFG, mode, B=1, device="cpu")
        what = "decoder forward + autograd backward" if mode == "train" else "decoder forward + heads"
        return step, "reference", f"unmodified reference TransformerDecoder (baseline/_ref, fp32, eval, torch-CPU {n} threads), {what}"
    cfg, W, inp, lw = _host_case()
    if mode == "train":
        from oracle import decoder_torch
        import torch
        torch.set_num_threads(os.cpu_count() or 1)
        return (lambda: decoder_torch.train_step(W, inp, lw, cfg["layers"])), "port", HOST_KIND["train"]
    from oracle import decoder_np

    def step():
        hs, cls_hs, refs = decoder_np.decoder_forward(W, inp["tgt"], inp["memory"], inp["mask"], inp["pos"],
                                                      inp["refpoints_unsigmoid"], inp["orig_res"], cfg["layers"])
        decoder_np.detr_heads(W, hs, cls_hs, refs)
    return step, "port", HOST_KIND["infer"]


def cpu_oracle_clips_per_s(mode, max_seconds=25.0, min_reps=1):
    step, kind, desc = host_step_fn(mode)
    step()                                   # warm-up (allocator, thread pool)
    times = []
    t_all = time.time()
    while len(times) < min_reps or (time.time() - t_all < max_seconds and len(times) < 10):
        t0 = time.time()
        step()
        times.append(time.time() - t0)
    return 1.0 / float(np.median(times)), len(times), kind, desc


class _Work(dict):
    def __getitem__(self, mode):
        from oracle import synth
        c = synth.CONFIGS[CFG]
        shape = f"({c['layers']} layers, nq {c['nq']}, T' {c['tprime']}, S {c['h'] * c['w']}, K {c['K']}, F {c['F']}"
        if mode == "train":
            return (f"{CFG_NAME[CFG]} class-query decoder fwd + bwd {shape}; gradients of all parameters, memory, tgt, refpoints; "
                    f"nn.Dropout(p={DROPOUT_P}) on at the residual-branch / FFN-hidden sites)")
        return f"{CFG_NAME[CFG]} class-query decoder forward + heads {shape})"


WORK = _Work()
HOST_KIND = {"train": "fp32 torch-CPU restatement of the reference decoder (oracle/decoder_torch.py), autograd backward",
             "infer": "fp32 numpy oracle port (oracle/decoder_np.py)"}


def bench_config(mode, B, world):
    """The `config` object of the JSON line -- identical for both arms (the reference arm measures a bounded sample of it)."""
    return {"workload": WORK[mode] + f", {B} clips/GPU", "mode": mode, "config": CFG, "batch_per_gpu": B,
            "parallelism": f"clip-sharded x{world}" + (", one NCCL all-reduce of the flat fp32 gradient per step" if mode == "train" and world > 1 else ""),
            "l2": "4 input sets cycled (128 MB) + GBs of intermediates per step >> 126 MB L2"}


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path (the unmodified reference decoder from
    baseline/_ref), all host threads, EXACTLY --warmup + --steps steps; each step is a bounded sample of the workload: ONE clip
    (our arm's step is `batch_per_gpu` clips), stated in cpu_baseline.sample.  Rank 0 only."""
    if rank != 0:
        return
    step, kind, desc = host_step_fn(args.mode)
    for _ in range(args.warmup):
        step()
    t0 = time.time()
    for _ in range(args.steps):
        step()
    dt = time.time() - t0
    v = args.steps / dt
    cores = os.cpu_count()
    sample = (f"{args.steps} timed steps x 1 clip per step (bounded sample: the GPU arm's step is {args.batch} clips/GPU); {desc}; "
              f"wall clock, {args.warmup} warm-up steps")
    print(json.dumps({
        "impl": "reference", "metric": "decoder clips/s", "value": v, "unit": "clips/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args.mode, args.batch, max(args.gpus, 1)),
        "clips_per_step": 1,
        "cpu_baseline": {"value": v, "unit": "clips/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=0, help="clips per GPU (default: 32 for the AVA configs, 1 for UCF / JHMDB)")
    ap.add_argument("--config", default="ava_vitb", choices=sorted(FWD_GFLOP))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="train", choices=["train", "infer"])
    ap.add_argument("--dropout", type=float, default=0.1, help="decoder nn.Dropout p in the training step (configuration/*.yaml DROPOUT 0.1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the other BASELINE configs / full training step extra keys")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    global CFG, DROPOUT_P
    CFG = args.config
    DROPOUT_P = args.dropout
    if args.batch <= 0:
        args.batch = DEFAULT_BATCH[CFG]

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from class_query_vad_b200 import DecoderEngine, _lib
    from oracle import synth

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    cfg = synth.CONFIGS[CFG]
    B = args.batch
    BT = B * cfg["tprime"]
    nq, K, Lr, F = cfg["nq"], cfg["K"], cfg["layers"], cfg["F"]
    W = synth.make_decoder_weights(K, Lr, F, seed=0)
    eng = DecoderEngine(W, nq=nq, K=K, layers=Lr, F=F, dtype=torch.bfloat16, device=dev, out_f32=False)
    lib = _lib.lib()

    # synthetic inputs: NSETS different batches cycled so that no step finds its inputs in L2 (126 MB); the
    # intermediates of one step (~0.8 GB) exceed L2 by themselves.
    NSETS = 4
    host_sets, dev_sets = [], []
    for s in range(NSETS):
        inp = synth.make_decoder_inputs(CFG, B, seed=100 * rank + s)
        hp = {k: torch.from_numpy(np.ascontiguousarray(inp[k])).pin_memory() for k in ("tgt", "memory", "pos", "mask", "refpoints_unsigmoid")}
        host_sets.append(hp)
        dev_sets.append({k: v.to(dev) for k, v in hp.items()})
    orig_res = (cfg["h"], cfg["w"])
    det_local = torch.empty((BT, nq, K + 4 + 3), dtype=torch.float32, device=dev)
    det_all = torch.empty((world * BT, nq, K + 4 + 3), dtype=torch.float32, device=dev) if world > 1 else det_local
    det_host = torch.empty((BT, nq, K + 4 + 3), dtype=torch.float32).pin_memory()
    from class_query_vad_b200.dist import allreduce_gradients
    OVERLAP = os.environ.get("CQVAD_BENCH_OVERLAP") is not None
    if world > 1 and args.mode == "train" and OVERLAP:
        eng._grad_table()
        eng.enable_layer_events()
    lw = synth.make_loss_weights(cfg, B, seed=1)
    g_hs = torch.from_numpy(lw["w_hs"]).to(dev).bfloat16()
    g_cls = torch.from_numpy(lw["w_cls"]).to(dev).bfloat16()
    g_refs = torch.from_numpy(lw["w_refs"]).to(dev)
    res_host = torch.empty((Lr * BT * nq * 4 + BT * nq * 256,), dtype=torch.float32).pin_memory()
    launches = {"n": 0}

    def infer_step(inp):
        out = eng.forward(inp["tgt"], inp["memory"], inp["mask"], inp["pos"], inp["refpoints_unsigmoid"], orig_res,
                          heads=True, skip_cls_hs=False)
        # detections of the last layer: [B, nq, K | 4 | 3] (the row format of utils/video_action_recognition.py:234)
        torch.cat([out["pred_logits"][-1], out["pred_boxes"][-1], out["pred_logits_b"][-1]], dim=-1, out=det_local)
        if world > 1:
            dist.all_gather_into_tensor(det_all, det_local)
        launches["n"] = eng.last_launches + 1 + (1 if world > 1 else 0)
        return out

    step_no = {"n": 0}

    def train_step(inp):
        step_no["n"] += 1          # nn.Dropout(0.1) ON, as in the reference training loop: a fresh Philox stream per step
        out = eng.forward_train(inp["tgt"], inp["memory"], inp["mask"], inp["pos"], inp["refpoints_unsigmoid"], orig_res,
                                dropout_p=args.dropout, seed=1000 + step_no["n"])
        eng.backward(g_hs, g_cls, g_refs, zero=True, named=False)
        if world > 1:       # one NCCL all-reduce of the flat gradient after the backward.  CQVAD_BENCH_OVERLAP=1: per-layer buckets on a
            # communication stream released by the layers' gradient-complete events -- measured no faster at 2 and 8 GPUs (the
            # NCCL CTAs contend with the persistent one-CTA-per-SM kernels of the backward): DESIGN.md section 8
            allreduce_gradients(eng, overlap=OVERLAP)
        launches["n"] = eng.last_launches + eng.last_launches_bwd + (1 if world > 1 else 0)
        return out

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def read_profile(steps):
        prof = {}
        for c in range(lib.cqvad_profile_num_classes()):
            tot, sc, ln = ctypes.c_double(), ctypes.c_long(), ctypes.c_long()
            lib.cqvad_profile_read(c, ctypes.byref(tot), ctypes.byref(sc), ctypes.byref(ln))
            if sc.value:
                fl, by = ctypes.c_double(), ctypes.c_double()
                lib.cqvad_profile_read_work(c, ctypes.byref(fl), ctypes.byref(by))
                prof[lib.cqvad_profile_class_name(c).decode()] = dict(ms_per_step=tot.value / steps, scopes=sc.value // max(steps, 1),
                                                                     launches=ln.value // max(steps, 1), flops_per_step=fl.value / steps,
                                                                     bytes_per_step=by.value / steps)
        return prof

    def measure(step, steps, result_to_host):
        """(device-resident ms, profile, launches per step, e2e ms) of `step`."""
        for i in range(args.warmup):
            step(dev_sets[i % NSETS])
        sync_all()
        ms = timed(lambda i: step(dev_sets[i % NSETS]), steps)
        n_launch = launches["n"]
        # per-kernel-class breakdown (and the roofline kernel's launch time) from a separate short pass: the event scopes
        # cost a few % and must not sit inside the timed region
        psteps = min(steps, 5)
        lib.cqvad_profile_enable(1)
        timed(lambda i: step(dev_sets[i % NSETS]), psteps)
        prof = read_profile(psteps)
        lib.cqvad_profile_enable(0)

        # end to end: every step copies ITS inputs from pinned host memory and reads its result back, inside the timed region.
        # The copies are double-buffered the way a data loader would do it: step i+1's H2D runs on a copy stream while step i
        # computes (one event per step orders them); the first step's copy is exposed.
        copy_stream = torch.cuda.Stream(device=dev)
        pending = {}
        bufs = [{k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in host_sets[0].items()} for _ in range(2)]
        done = [None, None]          # "the step that used this buffer has finished" events

        def stage(i):
            b = i % 2
            with torch.cuda.stream(copy_stream):
                if done[b] is not None:
                    copy_stream.wait_event(done[b])
                for k, v in host_sets[i % NSETS].items():
                    bufs[b][k].copy_(v, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return bufs[b], ev

        def e2e_step(i):
            inp, ev = pending.pop(i) if i in pending else stage(i)
            cur = torch.cuda.current_stream()
            cur.wait_event(ev)
            pending[i + 1] = stage(i + 1)
            out = step(inp)
            result_to_host(out)
            d = torch.cuda.Event()
            d.record(cur)
            done[i % 2] = d
        for i in range(2):
            e2e_step(i)
        pending.clear()
        ms_e2e = timed(e2e_step, steps)
        pending.clear()
        return ms, prof, n_launch, ms_e2e

    def train_result(out):
        n1 = Lr * BT * nq * 4
        res_host[:n1].copy_(out["refs"].reshape(-1), non_blocking=True)
        res_host[n1:].copy_(out["hs"][-1].reshape(-1).float(), non_blocking=True)

    def infer_result(out):
        det_host.copy_(det_local, non_blocking=True)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    h2d = sum(host_sets[0][k].numel() * host_sets[0][k].element_size() for k in host_sets[0])
    main_mode = args.mode
    if main_mode == "train":
        ms, prof, n_launch, ms_e2e = measure(train_step, args.steps, train_result)
        d2h = res_host.numel() * 4
        isteps = max(3, min(args.steps, 10))
        ims, iprof, in_launch, ims_e2e = measure(infer_step, isteps, infer_result)
    else:
        ms, prof, n_launch, ms_e2e = measure(infer_step, args.steps, infer_result)
        d2h = det_host.numel() * 4
    clocks = sampler.stop() if rank == 0 else None
    extra = {}
    if main_mode == "train" and CFG == "ava_vitb" and not args.no_extras:
        # the other BASELINE.json configurations, at this N (every rank takes part: weak scaling + the gradient all-reduce)
        del eng, dev_sets
        torch.cuda.empty_cache()
        for key, fn in (("other_configs", lambda: other_config_inference(dev, world, timed)),
                        ("full_training_step", lambda: full_training_step(dev, world))):
            try:
                extra[key] = fn()
            except Exception as e:                  # the headline never depends on the extra measurements
                extra[key] = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
                if world > 1:
                    raise

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    clips = world * B * args.steps
    value = clips / (ms / 1e3)
    peaks = load_peaks()
    # ---- per-class rooflines (every class whose kernels report algorithmic work) and the DOMINANT class = largest ms/step ----
    # times: CUDA events around each kernel on the stream it really runs on (the weight-gradient stream included), from the
    # short profile pass that follows the timed region; FLOPs / bytes: algorithmic (operands once, result once)
    ridge = peaks["tf_sust"] * 1e12 / (peaks["hbm"] * 1e9)
    classes = {}
    for k, v in prof.items():
        if v.get("flops_per_step", 0) <= 0:
            continue
        sec = v["ms_per_step"] * 1e-3
        tf, gbs = v["flops_per_step"] / sec / 1e12, v["bytes_per_step"] / sec / 1e9
        classes[k] = {"ms_per_step": round(v["ms_per_step"], 4), "launches": v["scopes"], "gflop": round(v["flops_per_step"] / 1e9, 1),
                      "mbytes": round(v["bytes_per_step"] / 1e6, 1), "tflops": round(tf, 1), "frac_tensor": round(tf / peaks["tf_sust"], 3),
                      "gbs": round(gbs, 1), "frac_hbm": round(gbs / peaks["hbm"], 3),
                      "bound": "tensor" if v["flops_per_step"] / max(v["bytes_per_step"], 1) > ridge else "hbm"}
    traffic_tab = {}
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        traffic_tab = json.load(open(tp))
    if classes:
        # candidates: classes that carry at least 2 % of the step's algorithmic bytes or FLOPs.  The small-row classes (localisation
        # chain, [480 x C] operands, 0.2 % of the work) run on their own stream UNDER the large class-branch kernels: their in-situ
        # time is mostly queueing behind persistent 148-CTA kernels, not work on the critical path, so they never bound the step.
        tot_b = sum(c_["mbytes"] for c_ in classes.values()) or 1.0
        tot_f = sum(c_["gflop"] for c_ in classes.values()) or 1.0
        for c_ in classes.values():
            c_["share_of_step_work"] = round(max(c_["mbytes"] / tot_b, c_["gflop"] / tot_f), 4)
        cand = [k for k in classes if classes[k]["share_of_step_work"] >= 0.02] or list(classes)
        dom = max(cand, key=lambda k: classes[k]["ms_per_step"])
        c = classes[dom]
        per_launch_ms = c["ms_per_step"] / max(c["launches"], 1)
        if c["bound"] == "tensor":
            roof = {"bound": "tensor", "kernel": dom, "achieved": c["tflops"], "peak": peaks["tf_sust"], "unit": "TFLOP/s",
                    "frac": c["frac_tensor"], "peak_source": peaks["src"] + " sustained bf16"}
        else:
            roof = {"bound": "hbm", "kernel": dom, "achieved": c["gbs"], "peak": peaks["hbm"], "unit": "GB/s", "frac": c["frac_hbm"],
                    "peak_source": peaks["src"] + " HBM copy"}
        roof.update({"traffic": (traffic_tab.get(dom) or {}).get("dram_bytes_per_launch"), "launch_ms": round(per_launch_ms, 5),
                     "launches_per_step": c["launches"], "ms_per_step": c["ms_per_step"],
                     "selection": "largest ms/step among the kernel classes carrying >= 2 % of the step's algorithmic bytes or FLOPs (see kernel_classes)",
                     "timing": "in situ: CUDA events on each kernel's own stream, concurrent streams enabled"})
    else:
        roof = None
    # the best kernel of the library, for reference (round 1 reported this one as the roofline kernel)
    S = cfg["h"] * cfg["w"]
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        v, reps, kind, desc = cpu_oracle_clips_per_s(main_mode)
        cpu = {"value": v, "unit": "clips/s", "cores": os.cpu_count(), "kind": kind,
               "sample": f"median of {reps} x 1 clip, same workload shape ({WORK[main_mode]}); {desc}"}
    gflop = 3.0 * FWD_GFLOP[CFG] if main_mode == "train" else FWD_GFLOP[CFG]
    line = {
        "metric": "decoder clips/s", "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": bench_config(main_mode, B, world),
        "decoder_tflops_per_gpu": value * gflop / 1e3 / world,
        "clocks": clocks,
        "e2e": {"value": clips / (ms_e2e / 1e3), "unit": "clips/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(n_launch * args.steps),
        "roofline": roof,
        "cpu_baseline": cpu,
        "breakdown_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in prof.items()},
        "kernel_classes": classes,
    }
    if main_mode == "train":
        iclips = world * B * isteps
        line["inference_forward"] = {
            "workload": WORK["infer"], "value": iclips / (ims / 1e3), "unit": "clips/s", "ms_per_step": ims / isteps, "steps": isteps,
            "e2e": iclips / (ims_e2e / 1e3), "gpu_launches_per_step": int(in_launch),
            "decoder_tflops": iclips / (ims / 1e3) * FWD_GFLOP[CFG] / 1e3 / world,
            "breakdown_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in iprof.items()}}
    line.update(extra)
    if world == 1:
        try:
            line["reference_gpu_eager"] = reference_gpu_eager(dev, main_mode, B)
        except Exception as e:
            line["reference_gpu_eager"] = {"error": str(e)[:200]}
    if main_mode == "train" and world == 1 and CFG == "ava_vitb":
        try:
            line["encoder_layer"] = encoder_layer_numbers(dev)
        except Exception as e:                      # the headline never depends on the extra measurement
            line["encoder_layer"] = {"error": str(e)[:200]}
        try:
            line["transformer_forward"] = transformer_forward_numbers(dev)
        except Exception as e:
            line["transformer_forward"] = {"error": str(e)[:200]}
        try:
            line["latency_batch1"] = batch1_latency(dev)
        except Exception as e:
            line["latency_batch1"] = {"error": str(e)[:200]}
        try:
            line["vit_neck"] = vit_neck_numbers(dev)
        except Exception as e:
            line["vit_neck"] = {"error": str(e)[:200]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def full_training_step(dev, world, B=2, steps=4, warmup=2):
    """BASELINE.json configs[4] (AVA22_ViT-B-train.yaml: batch 2 per GPU, 6 + 6 layers, 8 sampling points, single-frame decoding):
    ONE optimizer step through the public modules -- pinned-host pyramid -> H2D -> drop-in `Transformer` (level flatten + level_embed,
    6 deformable encoder layers with the MSDA-3D ops path, inter-stage resample, 6-layer class-query decoder) -> DETR heads (Philox
    Dropout(0.5) ON) -> device Hungarian matcher + SetCriterionAVA -> backward of the whole chain -> NCCL all-reduce of the flat
    gradient -> fused clip_grad_norm_ + AdamW (train.py:126-167).  The modules run in train() mode: encoder / decoder nn.Dropout(0.1) at
    the residual-branch / FFN-hidden sites and Dropout(0.5) on the class tokens are applied (not the attention-probability dropout).  CUDA events, max over ranks; the D2H read of the loss is inside the timed region."""
    import torch
    import torch.distributed as dist
    from class_query_vad_b200 import Transformer, DETRHeads, SetCriterionAVA, HungarianMatcherAVA, FlatAdamW, PositionEmbeddingSine_3D
    from oracle import synth
    shapes = [(8, 56, 56), (8, 28, 28), (8, 14, 14), (8, 7, 7)]
    K, nq, F_, P, Lr = 80, 15, 2048, 8, 6
    rank = dist.get_rank() if world > 1 else 0
    torch.cuda.reset_peak_memory_stats(dev)
    torch.manual_seed(1234)                         # identical initial weights on every rank (DDP broadcast in the reference)
    tr = Transformer(num_queries=nq, num_encoder_layers=6, num_decoder_layers=Lr, dim_feedforward=F_, enc_n_points=P, num_classes=K,
                     temp_len=16)
    We = synth.make_encoder_layer_weights(F_, 4, P, seed=5)
    Wd = synth.make_decoder_weights(K, Lr, F_, seed=0)
    sd = {"level_embed": torch.randn(4, 256)}
    for l in range(6):
        sd.update({f"encoder.layers.{l}." + k: torch.from_numpy(v) for k, v in We.items()})
    sd.update({"decoder." + k: torch.from_numpy(v) for k, v in Wd.items() if not k.startswith("heads.")})
    tr.load_state_dict(sd, strict=True)
    heads = DETRHeads()
    heads.bbox_embed = tr.decoder.bbox_embed        # shared module (models/model.py:100-101)
    model = torch.nn.ModuleDict({"transformer": tr, "heads": heads}).to(dev).train()
    refpoint_embed = torch.nn.Parameter(torch.randn(nq, 1, 4, device=dev))
    import warnings
    warnings.filterwarnings("ignore", message=".*dropout.*")
    matcher = HungarianMatcherAVA(cost_class=12, cost_bbox=5, cost_giou=2)
    crit = SetCriterionAVA(10, K, nq, matcher, {"loss_ce": 10, "loss_bbox": 5, "loss_giou": 2, "loss_ce_b": 1}, 0.1, ["labels", "boxes"], "ava")
    named = [(n, p_) for n, p_ in model.named_parameters()] + [("refpoint_embed", refpoint_embed)]
    opt = FlatAdamW(named, lr=1e-4, weight_decay=1e-2, max_norm=1.0, module=model)
    n_params = opt.numel
    masks = [torch.zeros((B,) + s_, dtype=torch.bool, device=dev) for s_ in shapes]
    pe = PositionEmbeddingSine_3D(256, normalize=True)

    class _NT:
        def __init__(self, m):
            self.tensors, self.mask = m, m
    poss = [pe(_NT(m)).to(torch.bfloat16) for m in masks]
    rs = np.random.RandomState(77 + rank)
    NS = 2
    host = [[torch.from_numpy(rs.standard_normal((B, 256) + s_).astype(np.float32)).bfloat16().pin_memory() for s_ in shapes] for _ in range(NS)]
    n_t = [1 + (i % 3) for i in range(B)]
    tb = np.concatenate([rs.uniform(0.2, 0.8, (B, 3, 2)), rs.uniform(0.05, 0.5, (B, 3, 2))], -1).astype(np.float32)
    tl = (rs.rand(B, 3, K) < 0.04).astype(np.float32); tl[..., 0] = 1.0
    tgt_boxes, tgt_labels = torch.from_numpy(tb).to(dev), torch.from_numpy(tl).to(dev)
    n_tgt = torch.tensor(n_t, dtype=torch.int32, device=dev)
    loss_host = torch.zeros(16).pin_memory()
    h2d = sum(t_.numel() * t_.element_size() for t_ in host[0])

    def step(i):
        srcs = [t_.to(dev, non_blocking=True) for t_ in host[i % NS]]
        hs, cls_hs, refs = model["transformer"](srcs, masks, poss, refpoint_embed)
        out = model["heads"](hs, cls_hs, refs)
        last = {k: out[k] for k in ("pred_logits", "pred_boxes", "pred_logits_b")}
        losses, _, grads = crit.total_and_grads(last, tgt_boxes, tgt_labels, n_tgt)
        torch.autograd.backward([last["pred_logits"], last["pred_boxes"], last["pred_logits_b"]], list(grads))
        opt.allreduce()
        opt.step(grad_scale=1.0 / world)
        loss_host.copy_(losses, non_blocking=True)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
    for i in range(warmup):
        step(i)
    sync()
    first_loss = float(loss_host[4])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    h0 = time.perf_counter()
    for i in range(steps):
        step(i)
    host_ms = (time.perf_counter() - h0) * 1e3 / steps          # host time to ENQUEUE a step (no synchronisation inside)
    e1.record()
    sync()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item()) / steps
    peak_gb = torch.cuda.max_memory_allocated(dev) / 2**30
    res = {"workload": "AVA22_ViT-B-train.yaml optimizer step: Transformer (6 deformable encoder layers on the 33 320-token ViT-B/224 "
                       "pyramid, MSDA-3D ops path, 8 points; resample; 6-layer class-query decoder) + heads (Dropout 0.5 on) + device "
                       f"matcher/criterion + backward + gradient all-reduce + clip + AdamW, bf16, {B} clips/GPU (the yaml's batch size)",
           "value": world * B / ms * 1e3, "unit": "clips/s", "ms_per_step": round(ms, 3), "steps": steps, "warmup": warmup,
           "n_gpus": world, "batch_per_gpu": B, "parameters": int(n_params), "allreduce_bytes_per_step": int(4 * n_params) if world > 1 else 0,
           "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 64, "loss_first": round(first_loss, 4), "loss_last": round(float(loss_host[4]), 4),
           "peak_memory_gib": round(peak_gb, 2), "host_enqueue_ms_per_step": round(host_ms, 2), "timing": "CUDA events, max over ranks, host copies inside the timed region (e2e)"}
    del model, opt, host
    torch.cuda.empty_cache()
    return res


def other_config_inference(dev, world, timed_fn, steps=3):
    """BASELINE.json configs[2] (AVA22_CSN152 decoder, batch-sharded inference) and configs[3] (UCF_ViT-B 24 class queries,
    JHMDB_ViT-B 21 class queries): decoder forward + heads, bf16, inputs resident, clips sharded over the ranks (weak scaling)."""
    import torch
    from class_query_vad_b200 import DecoderEngine
    from oracle import synth
    out = {}
    for name in ("ava_csn152", "ucf_vitb", "jhmdb_vitb"):
        c = synth.CONFIGS[name]
        Bc = DEFAULT_BATCH[name]
        W = synth.make_decoder_weights(c["K"], c["layers"], c["F"], seed=0)
        eng = DecoderEngine(W, nq=c["nq"], K=c["K"], layers=c["layers"], F=c["F"], dtype=torch.bfloat16, device=dev, out_f32=False)
        sets = []
        for s_ in range(2):
            inp = synth.make_decoder_inputs(name, Bc, seed=50 + s_)
            sets.append({k: torch.from_numpy(np.ascontiguousarray(inp[k])).to(dev) for k in ("tgt", "memory", "pos", "mask", "refpoints_unsigmoid")})
        res = (c["h"], c["w"])
        fn = lambda i: eng.forward(sets[i % 2]["tgt"], sets[i % 2]["memory"], sets[i % 2]["mask"], sets[i % 2]["pos"],
                                   sets[i % 2]["refpoints_unsigmoid"], res, heads=True)
        for i in range(3):
            fn(i)
        ms = timed_fn(fn, steps) / steps
        frames = Bc * c["tprime"]
        out[name] = {"workload": f"{CFG_NAME[name]} class-query decoder forward + heads ({c['layers']} layers, nq {c['nq']}, T' {c['tprime']}, "
                                 f"S {c['h'] * c['w']}, K {c['K']}), {Bc} clips/GPU", "value": round(world * Bc / ms * 1e3, 2), "unit": "clips/s",
                     "frames_per_s": round(world * frames / ms * 1e3, 1), "ms_per_step": round(ms, 3), "n_gpus": world,
                     "decoder_tflops_per_gpu": round(Bc / ms * 1e3 * FWD_GFLOP[name] / 1e3, 1)}
        del eng, sets
        torch.cuda.empty_cache()
    return out


def vit_neck_numbers(dev, B=4, iters=5):
    """the ViT simple-feature-pyramid neck (4 levels) on BASELINE configs[0]'s features
    (8x14x14x768 per clip) -> the encoder's 33 320-token sequence, bf16, inputs resident, CUDA events, median."""
    import torch
    from class_query_vad_b200 import SimpleFeaturePyramid
    torch.manual_seed(3)
    neck = SimpleFeaturePyramid(768).to(dev)
    x = [torch.randn((B, 768, 8, 14, 14), device=dev).bfloat16() for _ in range(4)]
    ts = []
    for _ in range(iters + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        neck.forward_tokens(x)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts[2:]))
    # algorithmic FLOPs per clip: ConvT GEMMs + 1x1x1 + 3x3x3 convs (2 * rows * K * N)
    r = 8 * 14 * 14
    fl = 2.0 * (r * 768 * 1536 + 4 * r * 384 * 768 + 16 * r * 192 * 256 + r * 768 * 1536 + 4 * r * 384 * 256 + r * 768 * 256 + (r // 4) * 768 * 256
                + (16 + 4 + 1) * r * 6912 * 256 + (8 * 7 * 7) * 6912 * 256)
    return {"workload": f"ViT-B simple feature pyramid (lateral_convs x4) on 8x14x14x768 features -> 33 320 tokens, bf16, {B} clips",
            "ms": round(ms, 3), "clips_per_s": round(B / ms * 1e3, 1), "gflop_per_clip": round(fl / 1e9, 1),
            "tflops": round(B * fl / ms / 1e9, 1)}


def batch1_latency(dev, iters=50):
    """BASELINE configs[0]'s shape on the GPU: AVA22_ViT-B decoder forward + heads for ONE clip (the case the reference runs on the
    CPU).  Launch-bound (317 kernels): timed eagerly and as a CUDA-graph replay (DecoderEngine.capture_forward), inputs resident,
    and end to end (pinned-host inputs -> static buffers -> replay -> detections back on the host)."""
    import torch
    from class_query_vad_b200 import DecoderEngine
    from oracle import synth
    cfg = synth.CONFIGS["ava_vitb"]
    W = synth.make_decoder_weights(cfg["K"], cfg["layers"], cfg["F"], seed=0)
    eng = DecoderEngine(W, nq=cfg["nq"], K=cfg["K"], layers=cfg["layers"], F=cfg["F"], dtype=torch.bfloat16, device=dev, out_f32=False)
    inp = synth.make_decoder_inputs("ava_vitb", 1, seed=3)
    host = {k: torch.from_numpy(np.ascontiguousarray(inp[k])).pin_memory() for k in ("tgt", "memory", "pos", "mask", "refpoints_unsigmoid")}
    d = {k: v.to(dev) for k, v in host.items()}
    res = (cfg["h"], cfg["w"])
    eager = lambda: eng.forward(d["tgt"], d["memory"], d["mask"], d["pos"], d["refpoints_unsigmoid"], res, heads=True)
    g = eng.capture_forward(d["tgt"], d["memory"], d["mask"], d["pos"], d["refpoints_unsigmoid"], res, heads=True)
    ref = eager()
    out = g()
    torch.cuda.synchronize()
    same = all(torch.equal(ref[k], out[k]) for k in ("hs", "refs", "pred_logits", "pred_boxes"))
    det_host = torch.empty((1, cfg["nq"], cfg["K"] + 7), dtype=torch.float32).pin_memory()

    def e2e():
        o = g(**host)
        det_host.copy_(torch.cat([o["pred_logits"][-1], o["pred_boxes"][-1], o["pred_logits_b"][-1]], dim=-1), non_blocking=True)

    def t(fn):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters
    return {"workload": "AVA22_ViT-B class-query decoder forward + heads, ONE clip (BASELINE configs[0] shape), bf16",
            "eager_ms": round(t(eager), 4), "cuda_graph_ms": round(t(lambda: g()), 4), "cuda_graph_e2e_ms": round(t(e2e), 4),
            "graph_replay_bit_identical_to_eager": bool(same), "launches": int(eng.last_launches)}


def transformer_forward_numbers(dev, B=8, iters=5):
    """The whole inference path this library covers, end to end on one GPU: ViT features [B,768,8,14,14] x4 -> simple-feature-pyramid
    neck -> level flatten + level_embed -> 6 deformable encoder layers (MSDA-3D, 33 320 tokens/clip) -> resample -> 6-layer class-query
    decoder (bf16) through the drop-in modules under torch.no_grad(); inputs resident; CUDA events, median."""
    import torch
    from class_query_vad_b200 import Transformer, SimpleFeaturePyramid, PositionEmbeddingSine_3D
    from oracle import synth
    torch.manual_seed(5)
    shapes = [(8, 56, 56), (8, 28, 28), (8, 14, 14), (8, 7, 7)]
    tr = Transformer(num_queries=15, num_encoder_layers=6, num_decoder_layers=6, dim_feedforward=2048, enc_n_points=8, num_classes=80, temp_len=16)
    We = synth.make_encoder_layer_weights(2048, 4, 8, seed=5)
    Wd = synth.make_decoder_weights(80, 6, 2048, seed=0)
    sd = {"level_embed": torch.randn(4, 256)}
    for l in range(6):
        sd.update({f"encoder.layers.{l}." + k: torch.from_numpy(v) for k, v in We.items()})
    sd.update({"decoder." + k: torch.from_numpy(v) for k, v in Wd.items() if not k.startswith("heads.")})
    tr.load_state_dict(sd, strict=True)
    tr = tr.to(dev).eval()
    tr.decoder.out_dtype = torch.bfloat16
    neck = SimpleFeaturePyramid(768).to(dev).eval()
    feats = [torch.randn((B, 768, 8, 14, 14), device=dev).bfloat16() for _ in range(4)]
    masks = [torch.zeros((B,) + s_, dtype=torch.bool, device=dev) for s_ in shapes]
    pe = PositionEmbeddingSine_3D(256, normalize=True)

    class _NT:
        def __init__(self, m):
            self.tensors, self.mask = m, m
    poss = [pe(_NT(m)).to(torch.bfloat16) for m in masks]
    refpoint = torch.randn(15, 1, 4, device=dev)

    def run():
        with torch.no_grad():
            lv = neck(feats)                                   # space_forward's channel-first maps, as models/model.py consumes them
            return tr([lv[str(i)] for i in range(4)], masks, poss, refpoint)
    ts = {"all": [], "neck": []}
    for _ in range(iters + 2):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        with torch.no_grad():
            lv = neck(feats)
        e1.record()
        with torch.no_grad():
            tr([lv[str(i)] for i in range(4)], masks, poss, refpoint)
        e2.record()
        torch.cuda.synchronize()
        ts["neck"].append(e0.elapsed_time(e1)); ts["all"].append(e0.elapsed_time(e2))
    ms, ms_neck = float(np.median(ts["all"][2:])), float(np.median(ts["neck"][2:]))
    del tr, neck, feats
    torch.cuda.empty_cache()
    return {"workload": f"ViT-B/224 AVA inference path after the backbone: neck + Transformer (6 encoder layers with MSDA-3D, resample, 6 decoder "
                        f"layers), bf16, {B} clips", "ms": round(ms, 3), "clips_per_s": round(B / ms * 1e3, 1), "neck_ms": round(ms_neck, 3),
            "encoder_plus_decoder_ms": round(ms - ms_neck, 3)}


def reference_gpu_eager(dev, mode, B, iters=5):
    """the UNMODIFIED reference decoder (baseline/_ref) run eagerly by PyTorch (cuBLAS / cuDNN / ATen
    kernels) on the same B200, same workload (B clips per step): fp32 with TF32 disabled (the parity oracle's arithmetic) and
    bf16 autocast (the reference's fastest stock path).  CUDA events, median of `iters` after 2 warm-ups.  A reported baseline."""
    import torch
    from oracle import ref_runner
    if not ref_runner.available():
        return {"unavailable": "baseline/_ref not installed (python oracle/install_ref.py)"}
    tf = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    out = {"workload": WORK[mode] + f", {B} clips per step, unmodified reference modules, eager PyTorch on cuda"}
    try:
        for tag, ac in (("fp32_tf32_off", False), ("bf16_autocast", True)):
            try:
                step = ref_runner.step_fn(CFG, mode, B=B, device=dev, autocast_bf16=ac)
                ts = []
                for i in range(iters + 2):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    step()
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                ms = float(np.median(ts[2:]))
                out[tag] = {"ms_per_step": round(ms, 3), "clips_per_s": round(B / ms * 1e3, 2)}
                del step
            except Exception as e:          # e.g. out of memory at this batch in fp32
                out[tag] = {"error": str(e)[:160]}
            torch.cuda.empty_cache()
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf
    return out


def encoder_layer_numbers(dev, B=4, iters=5):
    """SURVEY.md section 8f row 1, reported next to the headline: one deformable encoder layer (forward, and forward + backward) on
    the ViT-B/224 pyramid (Len = 33 320 tokens per clip), bf16, inputs resident, CUDA events, median of `iters`."""
    import numpy as np
    import torch
    from class_query_vad_b200 import DeformableTransformerEncoderLayer, DeformableTransformerEncoder
    from oracle import synth
    shapes = [(8, 56, 56), (8, 28, 28), (8, 14, 14), (8, 7, 7)]
    F_, P = 2048, 8
    W = synth.make_encoder_layer_weights(F_, 4, P, seed=5)
    inp = synth.make_encoder_inputs(1, shapes, seed=5)
    layer = DeformableTransformerEncoderLayer(d_model=256, d_ffn=F_, n_levels=4, n_heads=8, n_points=P)
    layer.load_state_dict({k: torch.from_numpy(v) for k, v in W.items()}, strict=True)
    layer = layer.to(dev).eval()
    sh = torch.tensor(shapes, dtype=torch.int64, device=dev)
    ls = torch.cat((sh.new_zeros((1,)), sh.prod(1).cumsum(0)[:-1]))
    refp = DeformableTransformerEncoder.get_reference_points(sh, torch.ones((B, 4, 3), device=dev), dev)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev).bfloat16().repeat(B, 1, 1).contiguous()
    src, pos = t(inp["src"]), t(inp["pos"])
    ev = lambda: torch.cuda.Event(enable_timing=True)
    tf, tt = [], []
    for _ in range(iters + 1):
        e0, e1 = ev(), ev()
        e0.record()
        with torch.no_grad():
            layer(src, pos, refp, sh, ls, None)
        e1.record()
        torch.cuda.synchronize()
        tf.append(e0.elapsed_time(e1))
    srcg, posg = src.clone().requires_grad_(True), pos.clone().requires_grad_(True)
    go = torch.randn_like(src)
    for _ in range(iters + 1):
        for p_ in layer.parameters():
            p_.grad = None
        e0, e1 = ev(), ev()
        e0.record()
        layer(srcg, posg, refp, sh, ls, None).backward(go)
        e1.record()
        torch.cuda.synchronize()
        tt.append(e0.elapsed_time(e1))
    mf, mt = float(np.median(tf[1:])), float(np.median(tt[1:]))
    # the same layer in the fp32 parity mode (CUDA-core FFMA GEMMs, fp32 sampling): what the 1e-3 tests run, not a performance path
    f32 = []
    s32, p32 = src.float(), pos.float()
    for _ in range(3):
        e0, e1 = ev(), ev()
        e0.record()
        with torch.no_grad():
            layer(s32, p32, refp, sh, ls, None)
        e1.record()
        torch.cuda.synchronize()
        f32.append(e0.elapsed_time(e1))
    return {"forward_ms_fp32_parity_mode": round(float(np.median(f32[1:])), 3), "workload": f"one DeformableTransformerEncoderLayer, ViT-B/224 pyramid (33 320 tokens/clip), F 2048, 8 points, bf16, {B} clips",
            "forward_ms": round(mf, 3), "forward_clips_per_s": round(B / mf * 1e3, 1),
            "train_step_ms": round(mt, 3), "train_step_clips_per_s": round(B / mt * 1e3, 1), "gemm_gflop_per_clip_forward": 96.1}


if __name__ == "__main__":
    main()
