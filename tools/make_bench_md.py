"""profiles/r01_raw/*.json -> the head of profiles/r01_bench.md (headline, inference, per-kernel tables).
The hand-written sections (GE2E sweep, history) below '## Fused GE2E' are kept."""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = os.path.join(ROOT, "profiles", "r01_raw")


def last(name):
    return json.loads([l for l in open(os.path.join(R, name)) if l.startswith("{")][-1])


b1, b2, b8, ref, i1, i8 = [last(n) for n in ("bench_1gpu.json", "bench_2gpu.json", "bench_8gpu.json",
                                              "bench_reference_arm.json", "bench_infer_1gpu.json",
                                              "bench_infer_8gpu.json")]
x1 = b1.get("extra", {})
out = []
w = out.append
w("# Round 1 -- measured numbers (B200, CUDA events, `bench.py`; raw JSON lines in profiles/r01_raw/)\n")
w("Roofline denominators: MEASURED_PEAKS.json -- HBM copy 6534.5 GB/s, bf16 1650.6 TFLOP/s burst / 1395.7 sustained.")
w("All runs saw `sw_power_cap` (SM clock 1750-1965 MHz under load, sampled through NVML during the timed region); box-to-box")
w("spread on this pool is +-4 %. A/B decisions in this round were taken on the same box (`tools/ab.sh`, `tools/ab_env.sh`).\n")
w("## Headline: GE2E training step (64 spk x 15 utt x T~U[140,180], dropout on, clip 1.0, fused RAdam + Modified-Noam)\n")
w("| GPUs | steps/s (device-resident inputs) | ms/step | steps/s end-to-end (pinned host -> `Device_Prefetcher` -> step -> loss read back) | SM MHz (median under load) |")
w("|---|---|---|---|---|")
for n, b in ((1, b1), (2, b2), (8, b8)):
    w("| %d | %.1f | %.2f | %.1f | %s |" % (n, b["value"], b["ms_per_step"], b["e2e"]["value"], b["clocks"]["sm_mhz"]))
w("| CPU (oracle port of the reference train step, torch CPU ops, %d cores, same box) | %.4f | %.0f | -- | `bench.py --impl reference` |\n"
  % (ref["cpu_baseline"]["cores"], ref["value"], ref["ms_per_step"]))
w("Weak scaling (every rank owns 64 speakers, one 9.8 MB NCCL all-reduce per step, no other collective): 2 GPUs = %.2f x, "
  "8 GPUs = %.2f x the 1-GPU line (%.0f %% / %.0f %% efficiency; the three lines ran on different boxes)."
  % (b2["value"] / b1["value"], b8["value"] / b1["value"], 100 * b2["value"] / b1["value"] / 2,
     100 * b8["value"] / b1["value"] / 8))
r = b1["roofline"]
w("\nRoofline object of the 1-GPU line: `%s`, %s-bound, %.0f of %.1f %s (frac %.3f), %.3f ms per launch, %.1f %% of the step; "
  "DRAM traffic %.3f GB per launch (ncu) vs %.3f GB algorithmic.\n"
  % (r["kernel"], r["bound"], r["achieved"], r["peak"], r["unit"], r["frac"], r["launch_ms"], 100 * r["share_of_step"],
     (r.get("traffic") or 0) / 1e9, r["alg_bytes_per_launch"] / 1e9))
w("## Inference\n")
w("| metric | value |")
w("|---|---|")
w("| d-vectors/s, 960 x 160-frame slices, 1 GPU | %d (%.2f ms per batch) |"
  % (x1["dvectors_per_sec_160f_1gpu"], x1["infer_ms_per_960x160_batch"]))
e2e_note = ("fp16 192-frame windows over PCIe through `Device_Prefetcher`, slices cut and upcast inside the prenet load by "
            "`GE2E.embed_windows`, asynchronous d-vector read-back")
w("| utterances/s, 5 x 64-frame slices / 32 overlap, 1 GPU (`--workload infer`) | %d; end-to-end (%s; %d MB H2D per step): %d |"
  % (i1["value"], e2e_note, i1["e2e"]["h2d_bytes_per_step"] / 1e6, i1["e2e"]["value"]))
w("| utterances/s, same, 8 GPUs (utterance shards, no collective) | %d (%.2f ms per 4000-utterance step per GPU); end-to-end: %d |"
  % (i8["value"], i8["ms_per_step"], i8["e2e"]["value"]))
w("| tensor-pipe fraction of sustained bf16 peak, F_min accounting: inference / training | %.3f / %.3f |\n"
  % (x1["infer_tensor_frac_of_sustained"], x1["train_tensor_frac_of_sustained"]))


def table(bd, title, minms):
    w("## %s\n" % title)
    w("| kernel tag | ms | launches | algorithmic TFLOP/s | algorithmic GB/s |")
    w("|---|---|---|---|---|")
    tot, nl = 0.0, 0
    key = "ms_per_step" if "ms_per_step" in next(iter(bd.values())) else "ms"
    for k, v in sorted(bd.items(), key=lambda kv: -kv[1][key]):
        tot += v[key]
        nl += v["launches"]
        if v[key] >= minms:
            w("| %s | %.3f | %d | %.1f | %.1f |" % (k, v[key], v["launches"], v.get("tflops", 0), v.get("gbs", 0)))
    w("| (total of all tags) | %.3f | %d | | |\n" % (tot, nl))


table(b1.get("breakdown") or x1.get("breakdown"),
      "Per-kernel breakdown of one training step (T = %d, in-library CUDA-event profiler, ms per step)" % b1["profiled_T"], 0.08)
table(b1.get("infer_breakdown") or x1.get("infer_breakdown"), "Inference breakdown (960 x 160, one plane)", 0.009)
path = os.path.join(ROOT, "profiles", "r01_bench.md")
old = open(path).read()
tail = old[old.index("## Fused GE2E, N speakers"):]
open(path, "w").write("\n".join(out) + tail)
print("\n".join(out[:22]))
