#!/usr/bin/env python
"""bench.py -- the hot-path benchmark (BASELINE.json: bootstrapped gates/sec per GPU).

One step = one pass of the hot path over one batch: 2^16 independent bootstrapped gates (config[1]: NAND on even
steps, XNOR on odd steps; reference keyset parameters), per GPU.  `value` times rs_gate_batch with inputs resident
in HBM; `e2e` times the reference-facing host-buffer call rs_gate_batch_host (pinned host buffers, H2D + D2H inside
the timed region).  `roofline` is the blind-rotation kernel against the FP64 roofline (258.4 MFLOP per bootstrap,
SURVEY.md 8d) with the peak measured live by rs_fp64_peak (MEASURED_PEAKS.json carries no FP64 figure).
`cpu_baseline` is the oracle's CPU port of the TFHE algorithm timed on the host cores (a reported baseline).

  python bench.py [--gpus N --steps K --warmup W]            this repo's CUDA path
  python bench.py --impl reference [...]                     the CPU path only (oracle port; TFHE itself is not installable)
Under torchrun (N>1) every rank runs its own 2^16-gate batch (weak scaling), time = max over ranks.
"""
from __future__ import annotations

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

GATES_PER_STEP = 1 << 16
FLOP_PER_PBS = 258.4e6            # 350 x (22 x 26112 + 163840)
BSK_FOURIER_BYTES = 114_688_000   # streamed once per wave of CTAs
KSK_TILED_BYTES = 1024 * 9 * 7 * 352 * 4   # device keyswitch table (digit-0 rows dropped)
LWE_WIRE_BYTES = 351 * 4
WORKLOAD = "gate microbench: 2^16 independent bootstrapped gates per step per GPU (NAND even steps / XNOR odd steps), keyset n=350 N=1024 l=10 Bgbit=3 t=9"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gates", type=int, default=GATES_PER_STEP, help="gates per step per GPU (default 2^16, the BASELINE config)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target wall time of the CPU baseline sample")
    ap.add_argument("--no-extra", action="store_true", help="skip the seconds/image side measurement")
    ap.add_argument("--nets", default="mnist/sign1024x1,mnist/sign1024x2,mnist/sign1024x3,mnist/cnn_builder,mnist/relu1024x1,cifar/binarynet",
                    help="nets timed for the seconds/image side measurement (BASELINE configs 1, 3, 4, 5 + one ReLU net)")
    ap.add_argument("--images", type=int, default=3, help="timed images per net (median reported), after one untimed warm-up image")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); smax = float(r[1]); power.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [c for c, p in zip(sm, power) if p > 300] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "power_w_max": max(power) if power else None, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------- CPU arm
def host_threads() -> int:
    """All host cores this process may use (torchrun exports OMP_NUM_THREADS=1, so the OpenMP default is not it)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def probe_upstream_tfhe() -> dict:
    """look for an installed upstream TFHE (libtfhe-spqlios-fma) on this box.  If it were present
    the CPU arm could bind it; it is not part of the image (no network), so this records the probe and the arm stays the port."""
    import ctypes.util
    found = None
    for name in ("tfhe-spqlios-fma", "tfhe-spqlios-avx", "tfhe-nayuki-portable"):
        path = ctypes.util.find_library(name)
        if path:
            found = path
            break
    if not found:
        try:
            out = subprocess.run(["ldconfig", "-p"], capture_output=True, text=True, timeout=10).stdout
            hits = [l.split("=>")[-1].strip() for l in out.splitlines() if "tfhe" in l.lower()]
            found = hits[0] if hits else None
        except Exception:
            found = None
    return {"libtfhe_found": found,
            "note": ("upstream TFHE is installed; the CPU arm still times the oracle port (no binding written: its key/ciphertext "
                     "layouts are unverifiable here)") if found else
                    "upstream TFHE v1.1 (libtfhe-spqlios-fma) is not installed on this box; the CPU arm is the oracle port"}


def cpu_baseline(ks, a, b, target_seconds: float, op: str = "NAND", check_against=None) -> dict:
    """Oracle port of the TFHE gate bootstrap on the host cores (bounded sample of the same workload).  `op` is the gate the GPU
    ran LAST, so the sample's ciphertexts can be compared with the timed batch's (sample_bit_exact_vs_gpu)."""
    from oracle import oracle as O
    oks = O.KeySet(ks.lwe_key, ks.tlwe_key, ks.bsk, ks.ksk)
    _ = oks.bsk_fft
    threads = host_threads()
    mu = 1 << 29
    t0 = time.perf_counter()
    O.gate(op, a[: 2 * threads], b[: 2 * threads], mu, oks, threads=threads)
    probe = time.perf_counter() - t0
    n = int(max(4 * threads, min(a.shape[0], 2 * threads * target_seconds / max(probe, 1e-3))))
    n = (n // threads) * threads
    t0 = time.perf_counter()
    out = O.gate(op, a[:n], b[:n], mu, oks, threads=threads)
    dt = time.perf_counter() - t0
    t1 = time.perf_counter()
    O.gate(op, a[:2], b[:2], mu, oks, threads=1)
    single = (time.perf_counter() - t1) / 2
    res = {"value": n / dt, "unit": "gates/s", "cores": threads, "kind": "port",
           "sample": f"{n} {op} gates of the step's batch, oracle FFT port (not the upstream TFHE binary), omp parallel for over gates",
           "ms_per_gate_single_core": single * 1e3, "upstream_probe": probe_upstream_tfhe()}
    if check_against is not None:
        res["sample_bit_exact_vs_gpu"] = bool(np.array_equal(out, check_against[:n]))
        res["sample_checked_gates"] = int(n)
    return res, n / dt


def _oracle_bits(O, bits, lwe_key, seed):
    mu = np.where(np.asarray(bits) == 1, 1 << 29, -(1 << 29))
    return O.encrypt(mu, 2.0 ** -25, lwe_key, seed)


def run_reference(args, rank: int):
    """--impl reference: the CPU implementation of the path (oracle port) with all host threads, bounded sample per step.
    Keygen and encryption are the oracle's too: this arm never loads libredsec_b200.so."""
    if rank != 0:
        return
    from oracle import oracle as O
    oks = O.keygen(0)
    _ = oks.bsk_fft
    threads = host_threads()
    rng = np.random.default_rng(1)
    n = max(2 * threads, 16)
    a = _oracle_bits(O, rng.integers(0, 2, n), oks.lwe_key, 11)
    b = _oracle_bits(O, rng.integers(0, 2, n), oks.lwe_key, 12)
    t0 = time.perf_counter(); O.gate("NAND", a, b, 1 << 29, oks, threads=threads); probe = time.perf_counter() - t0
    budget = 150.0 / max(args.steps + args.warmup, 1)               # whole run within a few minutes
    per_step = int(max(threads, min(4096, n * min(budget, 20.0) / max(probe, 1e-3))))
    per_step = max(threads, (per_step // threads) * threads)
    a = _oracle_bits(O, rng.integers(0, 2, per_step), oks.lwe_key, 13)
    b = _oracle_bits(O, rng.integers(0, 2, per_step), oks.lwe_key, 14)
    for s in range(args.warmup):
        O.gate("NAND" if s % 2 == 0 else "XNOR", a, b, 1 << 29, oks, threads=threads)
    t0 = time.perf_counter()
    for s in range(args.steps):
        O.gate("NAND" if s % 2 == 0 else "XNOR", a, b, 1 << 29, oks, threads=threads)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    sample = f"{per_step} gates per step (bounded sample of the 2^16-gate batch), oracle FFT port of the TFHE algorithm, {threads} threads"
    print(json.dumps({
        "impl": "reference", "metric": "bootstrapped_gates_per_sec", "value": value, "unit": "gates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample_gates_per_step": per_step,
                   "parity": "oracle parity unpinned vs upstream TFHE (no TFHE source or fixture in the reference tree)"},
        "cpu_baseline": {"value": value, "unit": "gates/s", "cores": threads, "kind": "port", "sample": sample,
                         "upstream_probe": probe_upstream_tfhe()},
        "e2e": {"value": value, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------- our arm
def run_ours(args, rank: int, local_rank: int, world: int):
    import torch
    import redsec_b200 as rs
    from redsec_b200 import client

    dist = None
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own banner ("NCCL version ...", printed to stdout when the environment
        # sets NCCL_DEBUG) goes to stderr instead
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)

    eng = rs.Engine(local_rank)
    stream = torch.cuda.Stream(device=dev)
    eng.set_stream(stream.cuda_stream)          # our kernels run on a stream torch events can see
    ks = client.keygen(0)
    eng.load_eval_key(ks.bsk, ks.ksk)
    fp64_peak = eng.fp64_peak_tflops()
    fp64_peak3 = eng.fp64_peak_three_operand_tflops()

    G = args.gates
    rng = np.random.default_rng(1 + rank)        # i.i.d. Bernoulli(1/2) inputs, alpha = 2^-25
    a_bits, b_bits = rng.integers(0, 2, G), rng.integers(0, 2, G)
    a = client.encrypt_bits(a_bits, ks.lwe_key, seed=100 + rank)
    b = client.encrypt_bits(b_bits, ks.lwe_key, seed=200 + rank)
    mu = client.EIGHTH
    d_a, d_b, d_out = eng.upload(a), eng.upload(b), eng.alloc(G)
    ops = ["NAND", "XNOR"]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident timing (`value`)
    for s in range(args.warmup):
        eng.gate(ops[s % 2], d_a, d_b, mu, d_out)
    eng.sync()
    sampler = ClockSampler(local_rank)
    eng.profile(True); eng.profile_reset()
    launches0 = eng.launch_count()
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for s in range(args.steps):
            eng.gate(ops[s % 2], d_a, d_b, mu, d_out)
        e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - launches0
    br_ms, br_n = eng.profile_get(rs.engine.K_BLIND_ROTATE)
    ks_ms, ks_n = eng.profile_get(rs.engine.K_KEYSWITCH)
    lin_ms, lin_n = eng.profile_get(rs.engine.K_LINEAR)
    eng.profile(False)
    last_op = ops[(args.steps - 1) % 2]
    out_dev = eng.download(d_out)

    # ---- correctness of what was timed: truth table of every gate of the last step (product decrypt)
    truth = (1 - (a_bits & b_bits)) if last_op == "NAND" else (1 - (a_bits ^ b_bits))
    verified = bool(np.array_equal(client.decrypt_bits(out_dev, ks.lwe_key), truth))

    # ---- end-to-end timing through the host-buffer C-ABI call (`e2e`): pinned host buffers, H2D + D2H in the region
    h_a = torch.from_numpy(a).pin_memory(); h_b = torch.from_numpy(b).pin_memory()
    h_out = torch.empty((G, 351), dtype=torch.int32).pin_memory()
    na, nb, nout = h_a.numpy(), h_b.numpy(), h_out.numpy()
    eng.gate_host(ops[0], na, nb, mu, nout)
    barrier()
    # rs_gate_batch_host is synchronous (returns when the result is in host memory), so the host clock between the two
    # barrier + device-sync points covers H2D + kernels + D2H of every step
    t0 = time.perf_counter()
    for s in range(args.steps):
        eng.gate_host(ops[s % 2], na, nb, mu, nout)
    barrier()
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_verified = bool(np.array_equal(nout.view(np.uint32), out_dev))

    # max over ranks (device time)
    t = torch.tensor([ms, e2e_wall_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max = float(t[0]), float(t[1])

    extra = {}
    if not args.no_extra:
        extra = nets_seconds_per_image(eng, ks, client, dist, rank, world, args.nets.split(","), args.images)

    cpu = None
    if rank == 0 and world >= 1:
        cpu, _ = cpu_baseline(ks, a, b, args.cpu_seconds, op=last_op, check_against=out_dev)

    if rank == 0:
        total_gates = world * G * args.steps
        value = total_gates / (ms_max * 1e-3)
        br_avg_s = br_ms / max(br_n, 1) * 1e-3
        achieved_tflops = FLOP_PER_PBS * G / br_avg_s / 1e12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r2_traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                if tj.get("gates_per_launch") == G:
                    traffic = tj.get("blind_rotate_dram_bytes_per_launch")
            except Exception:
                pass
        peaks = {}
        mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(mp):
            peaks = json.load(open(mp))
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        line = {
            "metric": "bootstrapped_gates_per_sec", "value": value, "unit": "gates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "gates_per_step_per_gpu": G, "per_gpu_gates_per_sec": value / world,
                       "l2": "inputs (2 x 92 MB of LWE rows) + 115 MB Fourier BSK + 104 MB KSK exceed the 126 MB L2; no explicit flush",
                       "keyset": "redsec_params_small_v2, product keygen seed 0", "verified_truth_table": verified,
                       "e2e_matches_device_path": e2e_verified},
            "e2e": {"value": total_gates / (e2e_ms_max * 1e-3), "unit": "gates/s", "h2d_bytes_per_step": 2 * G * LWE_WIRE_BYTES,
                    "d2h_bytes_per_step": G * LWE_WIRE_BYTES, "api": "rs_gate_batch_host (pinned host buffers)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "fp64", "achieved": achieved_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved_tflops / fp64_peak, "traffic": traffic,
                         "kernel": "blind_rotate_ws_kernel", "launch_ms": br_avg_s * 1e3, "launches": int(br_n),
                         "algorithmic_flop_per_launch": FLOP_PER_PBS * G,
                         "peak_source": "measured live by rs_fp64_peak: unrolled dependent-free DFMA loop (two register operands + immediate, 32 DFMA per trip), best of 12 / 16 / 32 warps per SM; MEASURED_PEAKS.json has no FP64 figure (nominal 64 DFMA/clk/SM x 148 x 1.965 GHz = 37.2)",
                         "peak_three_register_fma": fp64_peak3,
                         "frac_of_three_register_fma_peak": achieved_tflops / fp64_peak3,
                         "note": "a DFMA with three distinct register operands issues every 3 cycles instead of 2 (register-file bound, scripts/probes/fp64_probe2.cu); "
                                 "peak_three_register_fma is that rate measured live, the ceiling of the MAC and twiddle FMAs",
                         "traffic_source": "static: read from profiles/r2_traffic.json, one ncu capture of a 2^16-ciphertext launch (a bench run is never profiled)",
                         "traffic_note": "DRAM bytes of one 2^16-ciphertext blind-rotate launch; varies 70-220 GB between boxes: the CTAs of a wave finish ~1 % apart, the next wave's CTAs "
                                         "then walk the 114.7 MB Fourier key out of step and re-read it from DRAM (L2 hit rate 75-90 %). RS_WS_GATE=1 re-aligns every wave: 6.8 GB, "
                                         "99 % hit rate, 1.2 % slower (profiles/r2_traffic_ab.txt) -- the kernel is FP64-bound and HBM at most 4 % busy, so the gate is off",
                         "hbm": {"algorithmic_bytes_per_launch": BSK_FOURIER_BYTES + G * (352 * 4 + 1028 * 4),
                                 "achieved_gbs": (BSK_FOURIER_BYTES + G * (352 * 4 + 1028 * 4)) / br_avg_s / 1e9,
                                 "peak_gbs": hbm_peak, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
                         "keyswitch": {"kernel": "keyswitch_mma_kernel (tcgen05.mma kind::i8, one-hot x u8-limb key, exact) + transpose prep" if G >= 2048 else "keyswitch_tiled_kernel",
                                       "launch_ms": ks_ms / max(ks_n / 2, 1),
                                       "bound": "tensor (int8) / L2->smem key stream" if G >= 2048 else "shared-memory bandwidth",
                                       "tensor_ops_per_launch": 2.0 * G * 73728 * 1408,
                                       "achieved_tops": 2.0 * G * 73728 * 1408 / (ks_ms / max(ks_n / 2, 1) * 1e-3) / 1e12 if G >= 2048 else None,
                                       "peak_tops_nominal_int8_dense": 4500.0,
                                       "peak_note": "MEASURED_PEAKS.json has no int8 figure; nominal dense int8 4.5 POP/s (2x the bf16 figure, whose measured burst is %.0f TF/s)" % peaks.get("bf16_tflops", 0.0),
                                       "l2_to_smem_bytes_per_launch": (G // 256) * 6 * 2304 * 8192 if G >= 2048 else (G // 64) * KSK_TILED_BYTES,
                                       "note": "the one-hot GEMM executes 8x the additions of the gather formulation (digit 0..7 slots); the gather kernel needs 24.9 ms per 2^16"},
                         "step_share": {"blind_rotate_ms": br_ms / args.steps, "keyswitch_ms": ks_ms / args.steps,
                                        "linear_ms": lin_ms / args.steps}},
            "cpu_baseline": cpu,
        }
        if extra:
            line["extra"] = {"encrypted_inference": extra, "n_gpus": world,
                             "note": "seconds per image = median over timed images after one warm-up image; whole network in one library "
                                     "call (rs_net_run), layers sharded by output channel across the GPUs, NCCL all-gather between layers "
                                     "on the engine stream; device time, max over ranks"}
        print(json.dumps(line))
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


def nets_seconds_per_image(eng, ks, client, dist, rank: int, world: int, names, images: int = 3) -> dict:
    """Side measurement (BASELINE.json's second metric): encrypted inference seconds/image of the reference's nets on `world`
    GPUs.  The whole network is ONE library call (rs_net_run): activations resident, layers neuron-sharded by output channel,
    NCCL all-gather between layers issued by the library on the engine stream, no host synchronisation between layers.
    Per net: one untimed warm-up image (fills the caching allocator and NCCL's buffers), then `images` timed images, each
    bracketed by barrier + device sync on both sides and timed as the slowest rank's device time; the median is reported."""
    import torch
    out = {}
    dev = torch.device("cuda", eng.device)
    try:
        from redsec_b200 import netspec, nets
        for name in names:
            spec = netspec.NETS[name]()
            label, px = netspec.load_image_csv(spec["image"])
            ct = client.encrypt(netspec.map_pixels(spec, px) * client.UNIT, ks.lwe_key, client.SECALPHA, 7)
            net = nets.EncryptedNet(eng, spec)
            net.build_tables(rank, world)                     # weights/tables on the device before the timed window (as prep() does)
            d = eng.upload(ct)
            comm = net.comm_for((dist, rank, world) if dist is not None else None)

            def barrier():
                if dist is not None:
                    dist.barrier()
                torch.cuda.synchronize(dev)

            net.run_native(d, comm).free()                    # warm-up image
            eng.sync()
            times, res = [], None
            stream = torch.cuda.ExternalStream(eng.stream_handle(), device=dev)
            for _ in range(max(1, images)):
                if res is not None:
                    res.free()
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                res = net.run_native(d, comm)
                e1.record(stream)
                barrier()
                t = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=dev)
                if dist is not None:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                times.append(float(t[0]))
            scores = client.decrypt(eng.download(res), ks.lwe_key, 4096)
            res.free()
            dt = float(np.median(times))
            key = name.replace("/", "_")
            out[key] = {"s_per_image": dt, "s_per_image_all": times, "images_timed": len(times), "bootstraps": int(net.bootstraps()),
                        "bootstraps_per_sec": net.bootstraps() / dt, "argmax": int(np.argmax(scores)), "label": int(label),
                        "scores": [int(v) for v in scores]}
            net.close()
            eng.pool_trim()
    except Exception as e:   # the side measurement must never break the headline line
        out["error"] = str(e)[:300]
    return out


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
