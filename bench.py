#!/usr/bin/env python
"""bench.py — Hexa8 K+P assembly throughput (BASELINE.json metric) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

A "step" = one NIST.computeElements + CSRGenerator.updateCSR equivalent on the whole mesh:
U, dU, stateRef on the device -> CSR data, P, F, stateTemp on the device.
Workload (N=1): BoxGen 100x100x100 C3D8 (1M elements) linear elastic = BASELINE configs[1].
N>1: weak scaling, every rank owns a 100-plane slab (1M elements) of a (100 N)x100x100 box.

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (CUDA events, max over ranks);
`e2e` = the same pass through the host-facing call: pinned host U,dU -> device, assemble,
P, F and the CSR values back to the host, all inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (elType, material, props, (nX, nY, nZ) per GPU, dU recipe)
    "boxgen100_c3d8_linearelastic": ("C3D8", "linearelastic", [2.1e4, 0.22], (100, 100, 100), "normal1e-3"),
    "boxgen200x100x100_c3d8_vonmises": ("C3D8", "vonmises", [2.1e4, 0.22, 355.0, 1000.0, 200.0, 1400.0], (200, 100, 100), "shear"),
    "boxgen100_c3d8tl_neohookewa": ("C3D8TL", "neohookewa", [91304.34783, 100000.0], (100, 100, 100), "normal1e-2h"),
    "boxgen100x100x50_c3d20_linearelastic": ("C3D20", "linearelastic", [2.1e4, 0.22], (100, 100, 50), "normal1e-3"),
}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(workload, path):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/traffic.json)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        return json.load(open(p)).get(f"{workload}:{path}")
    return None


def algorithmic_bytes_per_element(nn, nGp, nState, nnz, nEl, nNode):
    """conn + coords + U,dU + state in/out + CSR values + P,F."""
    rho = nNode / nEl
    return 4 * nn + 24 * rho + 48 * rho + 2 * 8 * nGp * nState + 8 * nnz / nEl + 48 * rho


def make_inputs(kind, coords, n, l, seed=0):
    rng = np.random.default_rng(seed)
    nDof = 3 * coords.shape[0]
    if kind == "normal1e-3":
        dU = 1e-3 * rng.standard_normal(nDof)
    elif kind == "normal1e-2h":
        h = l[0] / n[0]
        dU = 1e-2 * h * rng.standard_normal(nDof)
    elif kind == "shear":  # gamma_xy(y) = gamma_max y / lY, ~50 % plastic Gauss points
        G = 2.1e4 / (2 * 1.22)
        gmax = 2.0 * 355.0 / (np.sqrt(3.0) * G)
        dU = 1e-6 * rng.standard_normal(nDof)
        y = coords[:, 1]
        dU[0::3] += 0.5 * gmax * y * y / l[1]
    else:
        raise ValueError(kind)
    return dU


class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe): NVML
    (nvidia_ml_py) every 5 ms, falling back to polling `nvidia-smi` when NVML is unavailable."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.samples = []  # (sm_mhz, power_w, reasons set)
        self.sm_max = None
        self.stop = False
        self.nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None
        self.t = threading.Thread(target=self.run, daemon=True)

    def _nvml_sample(self):
        n = self.nvml
        sm = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
        pw = n.nvmlDeviceGetPowerUsage(self.h) / 1000.0
        get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        mask = int(get(self.h))
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
        return sm, pw, {k for k, bit in names.items() if mask & bit}

    def run(self):
        while not self.stop:
            try:
                if self.nvml is not None:
                    self.samples.append(self._nvml_sample())
                    time.sleep(0.005)
                    continue
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    f = [x.strip() for x in out.split(",")]
                    self.sm_max = float(f[1])
                    reasons = {nme for nme, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]) if v.lower().startswith("active")}
                    self.samples.append((float(f[0]), float(f[2]), reasons))
            except Exception:
                pass
            time.sleep(0.02)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples"]}
        sm = sorted(s[0] for s in self.samples)
        reasons = set()
        for s in self.samples:
            reasons |= s[2]
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.sm_max, "power_w_max": max(s[1] for s in self.samples),
                "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi", "reasons": sorted(reasons)}


def cpu_reference_leg(wl, sample_n, steps, warmup):
    """The reference's CPU algorithm for this path (oracle port; /root/reference is Python and cannot
    travel to the GPU box) on a bounded sample of the same workload."""
    from oracle import cpu_baseline

    return cpu_baseline.run(wl, sample_n, steps, warmup)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="boxgen100_c3d8_linearelastic", choices=sorted(WORKLOADS))
    ap.add_argument("--n", type=int, nargs=3, default=None, help="override elements per GPU (nX nY nZ), for testing")
    ap.add_argument("--cpu-sample", type=int, default=32, help="edge length of the CPU-baseline sample box")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--generic", action="store_true", help="force the generic two-phase path")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    elType, material, props, n, recipe = WORKLOADS[args.workload]
    if args.n:
        n = tuple(args.n)
    warmup = max(args.warmup, 3)

    if args.impl == "reference":
        if rank != 0:
            return 0
        res = cpu_reference_leg(args.workload, args.cpu_sample, max(1, args.steps), max(1, min(args.warmup, 2)))
        line = {
            "impl": "reference", "metric": "Hexa8 K+P assembly throughput", "value": res["value"], "unit": "Melem/s", "n_gpus": args.gpus,
            "steps": res["steps"], "warmup": res["warmup"], "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": args.workload, "sample": res["sample"]},
            "cpu_baseline": {"value": res["value"], "unit": "Melem/s", "cores": res["cores"], "kind": res["kind"], "sample": res["sample"]},
            "e2e": {"value": res["value"], "unit": "Melem/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist

    from edelweissfe_b200 import _lib
    from edelweissfe_b200.partition import SlabAssembly

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- the rank's slab of the (n[0]*world) x n[1] x n[2] BoxGen box -------------------------
    # (contiguous element-plane blocks, each GPU owning its CSR rows; interface rows exchanged over NCCL)
    l = (float(n[0]), float(n[1]), float(n[2]))
    if "20" in elType:
        from edelweissfe_b200 import ElementAssembly, box_mesh

        assert world == 1, "C3D20 runs on the generic path, single GPU"
        coords, conn = box_mesh(n[0], n[1], n[2], lX=l[0], lY=l[1], lZ=l[2], elType=elType)
        asm = ElementAssembly(elType, conn, coords, material, props, device=dev)
        slab = None
    else:
        slab = SlabAssembly((n[0] * world, n[1], n[2]), (l[0] * world, l[1], l[2]), elType, material, props, rank, world, dev)
        asm = slab.asm
    coords = asm.coords.cpu().numpy()
    dU = make_inputs(recipe, coords, n, l, seed=rank)
    hU = torch.from_numpy(dU.copy()).pin_memory()
    hdU = torch.from_numpy(dU.copy()).pin_memory()
    asm.U.copy_(hU)
    asm.dU.copy_(hdU)
    flags = _lib.EWB_FLAG_FORCE_GENERIC if args.generic else 0
    indptr, indices = asm.csr_pattern()

    def step():
        if slab is not None:
            slab.assemble(flags)
        else:
            asm.assemble(flags)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for _ in range(warmup):
        step()
    asm.poll()
    launches0 = asm.launch_count()
    stream = torch.cuda.current_stream(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        barrier()
    ms = e0.elapsed_time(e1)
    launches = asm.launch_count() - launches0
    asm.poll()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    nEl_total = asm.nEl * world
    value = nEl_total / (ms_per_step * 1e-3) / 1e6

    # ---- end to end through the host-facing call (pinned host buffers) -------------------------
    e2e = None
    if not args.no_e2e:
        hP = torch.empty(asm.nDof, dtype=torch.float64).pin_memory()
        hF = torch.empty(asm.nDof, dtype=torch.float64).pin_memory()
        hK = torch.empty(asm.nnz, dtype=torch.float64).pin_memory()

        def e2e_step():
            asm.U.copy_(hU, non_blocking=True)
            asm.dU.copy_(hdU, non_blocking=True)
            step()
            hP.copy_(asm.P, non_blocking=True)
            hF.copy_(asm.F, non_blocking=True)
            hK.copy_(asm.csr_data, non_blocking=True)
            asm.poll()  # synchronises; raises CutbackRequest on material failure like the reference

        e2e_steps = max(2, min(args.steps, 5))
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / e2e_steps
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": nEl_total / dt / 1e6, "unit": "Melem/s", "h2d_bytes_per_step": int(2 * 8 * asm.nDof),
               "d2h_bytes_per_step": int(8 * (2 * asm.nDof + asm.nnz)), "ms_per_step": dt * 1e3,
               "note": "pinned host U,dU in; P, F and all CSR values out (host scipy/pardiso consumer, nonlinearimplicitstatic.py:451-454)"}

        # context only (not the headline): the same call when the matrix stays on the device for a device-side consumer
        def e2e_step_resident():
            asm.U.copy_(hU, non_blocking=True)
            asm.dU.copy_(hdU, non_blocking=True)
            step()
            hP.copy_(asm.P, non_blocking=True)
            hF.copy_(asm.F, non_blocking=True)
            asm.poll()

        e2e_step_resident()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step_resident()
        barrier()
        dtr = (time.perf_counter() - t0) / e2e_steps
        if world > 1:
            t = torch.tensor([dtr], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dtr = float(t.item())
        e2e["matrix_left_on_device"] = {"value": nEl_total / dtr / 1e6, "unit": "Melem/s", "ms_per_step": dtr * 1e3,
                                        "d2h_bytes_per_step": int(8 * 2 * asm.nDof),
                                        "note": "same call without the CSR-value copy (PCIe bound: %.2f GB per step)" % (8e-9 * asm.nnz)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = hbm_peak()
    balg = algorithmic_bytes_per_element(asm.nn, asm.nGp, asm.nState, asm.nnz, asm.nEl, asm.nNode)
    achieved = balg * asm.nEl / (ms_per_step * 1e-3) / 1e9  # GB/s per GPU (per-rank launch)
    fused = bool(asm.lib.ewb_plan_is_box(asm.plan)) and not args.generic and "20" not in elType
    line = {
        "metric": "Hexa8 K+P assembly throughput", "value": value, "unit": "Melem/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "elements_per_gpu": asm.nEl, "box_per_gpu": list(n), "nnz_per_gpu": asm.nnz, "dofs_per_gpu": asm.nDof,
                   "path": "fused-sweep" if fused else "generic-two-phase", "l2": "inputs+outputs (>3.6 GB/step) larger than the 126 MB L2",
                   "partition": "x-slabs of %d element planes per GPU, ghost-plane rows %s, %d B per interface"
                                % (n[0], "stored into the upper neighbour's memory by the sweep kernel (NVLink peer stores) + 4-byte status all-reduce"
                                   if slab.exchange == "peer" else "sent to the upper neighbour (NCCL P2P)", slab.interface_bytes)
                                if world > 1 else "single GPU"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": measured_traffic(args.workload, "fused-sweep" if fused else "generic-two-phase"),
                     "peak_source": peak_src, "algorithmic_bytes_per_element": balg,
                     "kernel": "sweepKernel (1 launch = 1 step)" if fused else "computeElementsVij+gatherResidual+updateCsr (3 launches = 1 step)"},
        "clocks": clocks.summary(),
        "gpu_launches": int(launches),
    }
    if e2e:
        line["e2e"] = e2e
    if not args.no_cpu:
        res = cpu_reference_leg(args.workload, args.cpu_sample, 20, 1)
        line["cpu_baseline"] = {"value": res["value"], "unit": "Melem/s", "cores": res["cores"], "kind": res["kind"], "sample": res["sample"]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
