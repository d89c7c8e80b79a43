"""profiles/r02_scaling.md from profiles/r02_bench_n*.json."""
import glob, json, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")
def load(name):
    p = os.path.join(P, name)
    return json.load(open(p)) if os.path.exists(p) else None
n1 = load("r02_bench_n1.json")
rows_w, rows_s = [], []
for n in (1, 2, 4, 8):
    d = n1 if n == 1 else load(f"r02_bench_n{n}_weak.json")
    if d:
        rows_w.append((n, d))
    d = load(f"r02_bench_n{n}_strong512.json")
    if d:
        rows_s.append((n, d))
out = ["# Round 2 scaling (one 8 x B200 box, one process per GPU, NCCL over NVLink / NVSwitch)\n",
       "Step = one training step of the fuser path (erank fwd/bwd + CMFuser fwd/bwd incl. the Block + all-reduce of the 13.66 MB fp32 "
       "parameter-gradient bucket on a side stream + the packed (2C+2)-float statistic all-reduce in the forward); `value` = clips of all "
       "ranks / max-over-ranks device time (CUDA events); `e2e` adds the pinned-host -> device copy of every step's inputs and the read-back "
       "of the statistic.\n",
       "## Weak scaling (B = 64 clips per GPU, BASELINE.json configs[1])\n",
       "| GPUs | clips/s | ms/step | speed-up | e2e clips/s | grad all-reduce alone (ms) |", "|---:|---:|---:|---:|---:|---:|"]
base = rows_w[0][1]["value"] if rows_w else None
for n, d in rows_w:
    ar = d.get("grad_allreduce", {}).get("ms")
    out.append(f"| {n} | {d['value']:.0f} | {d['ms_per_step']:.2f} | {d['value'] / base:.2f}x | {d['e2e']['value']:.0f} | "
               f"{'' if ar is None else f'{ar:.2f}'} |")
if rows_s:
    out += ["", "## Strong scaling (global batch 512 sharded over the ranks, BASELINE.json configs[2]; T = C = 512 bf16)\n",
            "| GPUs | clips per GPU | clips/s | ms/step | e2e clips/s |", "|---:|---:|---:|---:|---:|"]
    for n, d in rows_s:
        out.append(f"| {n} | {d['config']['B_per_gpu']} | {d['value']:.0f} | {d['ms_per_step']:.2f} | {d['e2e']['value']:.0f} |")
    out.append("\n(One GPU cannot hold the B = 512 working set of the eigensolver comfortably inside the bench's four rotating input sets, "
               "so the strong-scaling column starts where it was measured; per-GPU throughput at B = 64 is the weak-scaling row.)")
open(os.path.join(P, "r02_scaling.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
