"""Collect what tools/r02_first_gpu_call.sh left in gpurun_out/ into one markdown summary for profiles/:
    python tools/r02_report.py [gpurun_out] > profiles/r02a_summary.md
Tolerant of missing / truncated files (a variant whose process trapped simply shows up as 'no result')."""
import glob
import json
import os
import re
import sys

D = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"


def read(name):
    p = os.path.join(D, name)
    return open(p, errors="replace").read() if os.path.exists(p) else None


def json_lines(name):
    out = []
    for ln in (read(name) or "").splitlines():
        ln = ln.strip()
        if ln.startswith("{"):
            try:
                out.append(json.loads(ln))
            except ValueError:
                pass
    return out


def last_line(name, pat):
    hits = [ln for ln in (read(name) or "").splitlines() if re.search(pat, ln)]
    return hits[-1].strip() if hits else "no result"


def num(x, fmt):
    return format(x, fmt) if isinstance(x, (int, float)) else "n/a"


print("# Round-2 opening GPU call: summary\n")
print("## Test processes (exit codes)\n```\n" + (read("r02a_summary.txt") or "r02a_summary.txt missing").strip() + "\n```\n")
for log in sorted(glob.glob(os.path.join(D, "r02a_pytest_*.log"))):
    tail = [ln for ln in open(log, errors="replace").read().splitlines() if ln.strip()][-1:]
    print(f"* `{os.path.basename(log)}`: {tail[0] if tail else 'empty'}")

print("\n## Whole forward, B = 16, r256 (tools/time_forward.py; CUDA events)\n")
print("| variant | result |\n|---|---|")
for tag in ("default", "pair", "raywarp", "rayfold", "styleprep"):
    print(f"| {tag} | {last_line(f'r02a_time_forward_{tag}.log', r'^skip_noise_draws')} |")

print("\n## CIPS kernel alone (tools/time_kernels.py)\n")
for tag in ("default", "pair"):
    txt = read(f"r02a_time_kernels_{tag}.log")
    print(f"**{tag}**\n```\n" + ("\n".join(txt.strip().splitlines()[-6:]) if txt else "no result") + "\n```")

print("\n## Contract bench lines\n")
print("| run | value (img/s) | e2e | e2e_uint8 | roofline kernel | frac | frac_of_burst | variants |\n|---|---|---|---|---|---|---|---|")
for name in ("r02a_bench.json", "r02a_bench_variants.json"):
    for j in json_lines(name):
        r = j.get("roofline") or {}
        print(f"| {name} | {num(j.get('value'), '.1f')} | {num((j.get('e2e') or {}).get('value'), '.1f')} | "
              f"{num((j.get('e2e_uint8') or {}).get('value'), '.1f')} | {r.get('kernel')} | {num(r.get('frac'), '.3f')} | "
              f"{num(r.get('frac_of_burst'), '.3f')} | {(j.get('config') or {}).get('variants')} |")

print("\n## HBM-bound ops (tools/bench_disc_ops.py, tools/bench_optim.py)\n")
print("| op | shape / params | ms | GB/s | frac of HBM peak | note |\n|---|---|---|---|---|---|")
for name in ("r02a_disc_ops_default.jsonl", "r02a_disc_ops_blur_tma.jsonl", "r02a_optim.jsonl"):
    for j in json_lines(name):
        if "op" not in j:
            continue
        ms = j.get("ms", j.get("fused_ms"))
        gbs = j.get("gbs", j.get("fused_gbs"))
        note = f"torch {num(j.get('torch_ms'), '.3f')} ms, x{num(j.get('speedup'), '.2f')}" if "torch_ms" in j \
            else name.replace("r02a_", "").replace(".jsonl", "")
        print(f"| {j['op']} | {j.get('shape', j.get('params'))} | {num(ms, '.4f')} | {num(gbs, '.0f')} | {num(j.get('frac'), '.3f')} | {note} |")

print("\n## pi-GAN renderer (tools/time_pigan.py)\n")
print("| impl | img | batch | ms | img/s | TFLOP/s | frac of tensor peak |\n|---|---|---|---|---|---|---|")
for j in json_lines("r02a_time_pigan.jsonl"):
    print(f"| {j.get('impl')} | {j.get('img_size')} | {j.get('batch')} | {num(j.get('ms'), '.3f')} | {num(j.get('img_per_s'), '.1f')} | "
          f"{num(j.get('tflops'), '.1f')} | {num(j.get('frac_of_tensor_peak'), '.3f')} |")

print("\n## Train step (tools/bench_train_step.py)\n")
print("| file | images/s | ms/step | config | losses finite |\n|---|---|---|---|---|")
for p in sorted(glob.glob(os.path.join(D, "r02a_train_*.json"))):
    rows = json_lines(os.path.basename(p))
    for j in rows:
        c = j.get("config") or {}
        print(f"| {os.path.basename(p)} | {num(j.get('value'), '.2f')} | {num(j.get('ms_per_step'), '.1f')} | "
              f"config {c.get('baseline_config')} r{c.get('resolution')} b{c.get('batch_per_gpu')} optim={c.get('optim')} "
              f"cips={c.get('cips_backend')} film={c.get('film_backend')} tf32={c.get('tf32_autograd')} | {j.get('finite')} |")
    if not rows:
        err = (read(os.path.basename(p).replace(".json", ".err")) or "").strip().splitlines()
        print(f"| {os.path.basename(p)} | no result | | {err[-1][:120] if err else ''} | |")
