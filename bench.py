#!/usr/bin/env python
"""Benchmark of the sam_road tiled-inference hot path on B200 (contract: see the task brief / DESIGN.md).

    python bench.py --gpus 1 --steps 5 --warmup 3                 # this framework (CUDA, sm_100a)
    python bench.py --impl reference --steps 3 --warmup 1         # reference algorithm on host CPU cores
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], `toponet_vitb_512_cityscale`): one "step" is one pass of the hot
path over one INFER_BATCH_SIZE=64 batch of synthetic 512x512 RGB tiles per GPU: ViT-B encoder + naive
mask decoder + TopoNet on 256 keypoints x 16 neighbour pairs per tile.  Weights are seeded random
tensors with the reference's state_dict layout; data is synthetic (no network for datasets/ckpts).

Printed JSON (one line, rank 0):
  value     tiles/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e       same metric through the host-buffer C-ABI call (pinned host tiles in, results out)
  roofline  dominant kernel class: algorithmic FLOPs / CUDA-event duration vs MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (port of the reference algorithm) timed on this box's host cores
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOAD = "toponet_vitb_512_cityscale"
CONFIG = dict(SAM_VERSION="vit_b", PATCH_SIZE=512, USE_SAM_DECODER=False, ENCODER_LORA=False,
              TOPONET_VERSION="normal", NO_SAM=False, INFER_BATCH_SIZE=64)
POINTS_PER_TILE = 256
METRIC = "512x512 ViT-B tiles/sec"

# algorithmic FLOPs per tile: encoder 194.50 G + naive decoder 0.84 G, TopoNet
# 10.89 MFLOP per 16-pair sample + 65.5 kFLOP per keypoint
FLOP_PER_TILE = 195.34e9 + POINTS_PER_TILE * (10.89e6 + 65.5e3)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(tflops=float(p["bf16_tflops_sustained"]), tflops_burst=float(p["bf16_tflops"]),
                    hbm=float(p["hbm_gbs"]), source="measured (MEASURED_PEAKS.json, sustained)")
    return dict(tflops=1400.0, tflops_burst=1590.0, hbm=6650.0,
                source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: NVML every 10 ms when the binding
    is importable (nvidia-ml-py), else one nvidia-smi query per 150 ms."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None
        self.source = "nvidia-smi"
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self._max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._nvml = pynvml
            self.source = "nvml"
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n = self._nvml
        sm = float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM))
        try:
            watts = n.nvmlDeviceGetPowerUsage(self._h) / 1000.0
        except Exception:
            watts = float("nan")
        try:
            mask = n.nvmlDeviceGetCurrentClocksEventReasons(self._h)
        except Exception:
            mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        bits = [getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8),
                getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)]
        self.rows.append([str(sm), str(self._max), str(watts)] +
                         ["Active" if mask & b else "Not Active" for b in bits])

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits"], capture_output=True,
                                         text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([c.strip() for c in out.splitlines()[0].split(",")])
            except Exception:
                pass
            self._stop.wait(0.01 if self._nvml is not None else 0.15)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        def num(x):
            try:
                return float(x)
            except Exception:
                return None
        sm = sorted(v for v in (num(r[0]) for r in self.rows if r) if v is not None)
        reasons = set()
        for r in self.rows:
            for n, v in zip(self.NAMES, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        mx = max((v for v in (num(r[1]) for r in self.rows if len(r) > 1) if v is not None), default=None)
        pw = [v for v in (num(r[2]) for r in self.rows if len(r) > 2) if v is not None and v == v]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx,
                "sm_mhz_min": sm[0] if sm else None, "power_w_max": max(pw) if pw else None,
                "reasons": sorted(reasons), "samples": len(self.rows), "source": self.source}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle (CPU port of the reference algorithm) on host cores
# ------------------------------------------------------------------------------------------------
def time_cpu_oracle(n_tiles: int, steps: int, warmup: int):
    import torch
    from oracle import samroad_oracle as O          # the only place bench.py executes oracle/
    from sam_road_b200 import synth
    spec = O.ModelSpec.from_config(CONFIG)
    sd = synth.make_state_dict(CONFIG, seed=0)
    rgb = synth.make_tiles(n_tiles, 512, seed=11, dtype=torch.float32)
    pts, prs, val = synth.make_topo_inputs(n_tiles, 512, POINTS_PER_TILE, seed=12, ragged=False)
    # all the host threads the process can really use: the affinity mask / cgroup quota may be far
    # below os.cpu_count() on a shared box, and oversubscribed eager PyTorch is several times slower,
    # so probe a few thread counts on one tile and keep the fastest
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            avail = max(1, min(avail, int(int(quota) / int(period))))
    except Exception:
        pass
    cands = sorted({avail, min(avail, 64), min(avail, 32), min(avail, 16), min(avail, 8)}, reverse=True)
    best_t, cores = None, avail
    with torch.no_grad():
        for c in cands:
            torch.set_num_threads(c)
            O.image_encoder(O.normalize_rgb(rgb[:1]), sd, spec)          # warm
            t0 = time.perf_counter()
            O.image_encoder(O.normalize_rgb(rgb[:1]), sd, spec)
            dt = time.perf_counter() - t0
            if best_t is None or dt < best_t:
                best_t, cores = dt, c
    torch.set_num_threads(cores)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            _, feat = O.infer_masks_and_img_features(sd, spec, rgb)
            O.infer_toponet(sd, spec, feat, pts, prs, val)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    total = sum(times)
    return dict(value=n_tiles * len(times) / total, ms_per_step=1e3 * total / len(times), cores=cores,
                sample=f"{n_tiles} tiles of 512x512 + TopoNet ({POINTS_PER_TILE} keypoints x 16 pairs) "
                       f"per step, {len(times)} timed steps after {warmup} warm-up, fp32, "
                       f"torch.set_num_threads({cores})")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = time_cpu_oracle(n_tiles=args.ref_tiles, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "tiles/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "tiles_per_step": args.ref_tiles,
                   "points_per_tile": POINTS_PER_TILE, "pairs_per_point": 16,
                   "note": "reference algorithm (oracle port, fp32 PyTorch eager) on host CPU cores; "
                           "each step is a bounded sample of the workload"},
        "cpu_baseline": {"value": r["value"], "unit": "tiles/s", "cores": r["cores"], "kind": "port",
                         "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "tiles/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------
def run_native(args):
    import torch
    import torch.distributed as dist
    from sam_road_b200 import SAMRoad, _lib, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the native arm has no CPU fallback "
                         "(use --impl reference for the CPU reference arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    B, P, NP = args.batch, 512, POINTS_PER_TILE

    net = SAMRoad(CONFIG)
    net.load_state_dict(synth.make_state_dict(CONFIG, seed=0), strict=True)
    net.eval().to(dev)

    # R distinct resident input batches: R * 50 MB of uint8 tiles > L2 (126 MB); the step's own
    # activations (>1 GB) also exceed L2 many times over, so no explicit L2 flush is needed.
    R = 3
    tiles = [synth.make_tiles(B, P, seed=100 * rank + r).to(dev) for r in range(R)]
    topo_host = [synth.make_topo_inputs(B, P, NP, seed=100 * rank + r, ragged=False) for r in range(R)]
    topo = [[t.to(dev) for t in th] for th in topo_host]

    gather_sc = gather_ts = None
    if world > 1:
        gather_sc = torch.empty((world * B, P, P, 2), dtype=torch.float32, device=dev)
        gather_ts = torch.empty((world * B, NP, 16, 1), dtype=torch.float32, device=dev)

    def step(i):
        r = i % R
        scores, feat = net.infer_masks_and_img_features(tiles[r])
        ts = net.infer_toponet(feat, *topo[r])
        if world > 1:   # the path's exchange step: per-tile mask + topology scores
            dist.all_gather_into_tensor(gather_sc, scores)
            dist.all_gather_into_tensor(gather_ts, ts)
        return scores, ts

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    if args.debug_gemm_mode:
        lib.samroad_debug_disable_2cta_gemm(args.debug_gemm_mode)
    handle = net._handle(dev)
    _lib.check(lib.samroad_timing_enable(handle, 1), "timing_enable")
    lib.samroad_launch_count(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        e0.record()
        for i in range(args.steps):
            step(args.warmup + i)
        e1.record()
        barrier()
    launches = int(lib.samroad_launch_count(0))
    ms = e0.elapsed_time(e1)
    buf = C.create_string_buffer(1 << 16)
    _lib.check(lib.samroad_timing_read(handle, buf, len(buf)), "timing_read")
    kernels = json.loads(buf.value.decode())
    _lib.check(lib.samroad_timing_enable(handle, 0), "timing_disable")
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    value = world * B * args.steps / (ms / 1e3)

    # ---- e2e: the host-buffer C-ABI call, pinned host tiles in, results out, every step ----------
    h_tiles = [t.cpu().pin_memory() for t in tiles]
    h_topo = [[t.contiguous().pin_memory() for t in (th[0], th[1], th[2].view(torch.uint8))]
              for th in topo_host]
    h_scores = torch.empty((B, P, P, 2), dtype=torch.float32).pin_memory()
    h_emb = torch.empty((B, 256, P // 16, P // 16), dtype=torch.float32).pin_memory()
    h_ts = torch.empty((B, NP, 16), dtype=torch.float32).pin_memory()

    def e2e_step(i):
        r = i % R
        p, q, v = h_topo[r]
        _lib.check(lib.samroad_infer_batch_host(
            handle, h_tiles[r].data_ptr(), _lib.U8, B, p.data_ptr(), _lib.I64, q.data_ptr(), _lib.I64,
            v.data_ptr(), NP, NP, 16, h_scores.data_ptr(), h_emb.data_ptr(), h_ts.data_ptr()),
            "samroad_infer_batch_host")

    for i in range(max(1, min(2, args.warmup))):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(i)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = t.item()
    e2e_value = world * B * args.steps / e2e_s
    h2d = sum(x.numel() * x.element_size() for x in (h_tiles[0], *h_topo[0]))
    d2h = sum(x.numel() * x.element_size() for x in (h_scores, h_emb, h_ts))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel class ---------------------------------------------------
    # DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the kernel classes captured
    # with `ncu --set full` at this workload: profiles/r01_gemm_f16_v4_raw.csv, r01_gemm_resid_v4_raw.csv
    # (proj: r01_gemm_resid_v3_raw.csv), r01_att_v6_raw.csv (tools/gpu/profile_r01.sh).  Only valid for
    # the default batch of 64 tiles.
    ncu_traffic = {"gemm_mlp_lin1": 105.514240e6 + 347.093760e6, "gemm_qkv": 104.339456e6 + 245.466624e6,
                   "gemm_mlp_lin2": 614.097152e6 + 172.180992e6, "gemm_proj": 303.192576e6 + 143.959552e6,
                   "attention_global": 302.231040e6 + 80.478464e6, "attention_window": 302.097664e6 + 78.441216e6}
    peaks = load_peaks()
    gemm_like = {k: v for k, v in kernels.items() if v["flops"] > 0}
    dom = max(gemm_like, key=lambda k: gemm_like[k]["ms"]) if gemm_like else None
    roofline = None
    if dom:
        d = kernels[dom]
        per_launch_ms = d["ms"] / d["launches"]
        achieved = d["flops"] / d["launches"] / (per_launch_ms * 1e-3) / 1e12
        roofline = {"kernel": dom, "bound": "tensor", "achieved": achieved, "peak": peaks["tflops"],
                    "unit": "TFLOP/s", "frac": achieved / peaks["tflops"],
                    "traffic": ncu_traffic.get(dom) if args.batch == 64 else None,
                    "traffic_source": "ncu --set full, profiles/r01_*_raw.csv (bytes per launch)",
                    "peak_source": peaks["source"], "launches": d["launches"],
                    "avg_launch_ms": per_launch_ms,
                    "algorithmic_flops_per_launch": d["flops"] / d["launches"]}
    total_kernel_ms = sum(v["ms"] for v in kernels.values())
    shares = {k: {"ms_per_step": v["ms"] / args.steps, "share": v["ms"] / total_kernel_ms,
                  "tflops": (v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["ms"] > 0 else 0.0,
                  "gbs": (v["bytes"] / (v["ms"] * 1e-3) / 1e9) if v["ms"] > 0 else 0.0,
                  "launches_per_step": v["launches"] / args.steps}
              for k, v in sorted(kernels.items(), key=lambda kv: -kv[1]["ms"])}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = time_cpu_oracle(n_tiles=2, steps=3, warmup=1)
        cpu = {"value": r["value"], "unit": "tiles/s", "cores": r["cores"], "kind": "port",
               "sample": r["sample"]}

    line = {
        "metric": METRIC, "value": value, "unit": "tiles/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f16 operands / f32 accumulate",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "tiles_per_step_per_gpu": B, "patch_size": P,
                   "points_per_tile": NP, "pairs_per_point": 16, "input_dtype": "uint8",
                   "l2_policy": f"{R} rotating resident input batches (> L2) and >1 GB of "
                                "activations per step; no explicit flush",
                   "parallelism": f"tile-sharded dp{world}" +
                                  (" + all_gather(mask scores, topo scores)" if world > 1 else "")},
        "e2e": {"value": e2e_value, "unit": "tiles/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s / args.steps,
                "call": "samroad_infer_batch_host (pinned host uint8 tiles -> mask scores, "
                        "embeddings, topology scores on host)", "timer": "perf_counter"},
        "gpu_launches": launches,
        "clocks": clocks.summary(),
        "roofline": roofline,
        "path_tensor_frac": value / world * FLOP_PER_TILE / 1e12 / peaks["tflops"],
        "algorithmic_gflop_per_tile": FLOP_PER_TILE / 1e9,
        # BASELINE's third metric: sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active of the
        # dominant kernels from the committed ncu --set full captures (not measured in this run)
        "tensor_pipe_pct_ncu": {"gemm_qkv": 72.5, "gemm_mlp_lin1": 66.6, "gemm_mlp_lin2": 79.8,
                                "attention_global": 22.3, "attention_window": 19.2,
                                "source": "profiles/r01_ncu_summary_v4.md"},
        "kernels": shares,
        "cpu_baseline": cpu,
    }
    emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


class _JsonStdout:
    """Keep stdout clean for the single JSON line: libraries (NCCL prints its version banner on the
    first communicator) write to fd 1, so fd 1 is pointed at stderr for the run and the JSON line goes
    to the saved original stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.fd = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, line: str):
        os.write(self.fd, (line + "\n").encode())

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.fd, 1)
        os.close(self.fd)


_OUT = None


def emit(line: str):
    if _OUT is not None:
        _OUT.emit(line)
    else:
        print(line, flush=True)


def main():
    global _OUT
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=CONFIG["INFER_BATCH_SIZE"])
    ap.add_argument("--ref-tiles", type=int, default=4,
                    help="tiles per step of the CPU reference arm (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--debug-gemm-mode", type=int, default=0,
                    help="A/B only: samroad_debug_disable_2cta_gemm bit mask (16 = no snake traversal)")
    args = ap.parse_args()
    with _JsonStdout() as out:
        _OUT = out
        if args.impl == "reference":
            run_reference(args)
        else:
            run_native(args)


if __name__ == "__main__":
    main()
