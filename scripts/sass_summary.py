"""Counts the SASS mnemonics that show how each hot kernel is built (B200_PROFILING.md, "what proves a Blackwell-native
kernel"): UBLKCP = TMA bulk copy, UBLKPF = TMA L2 prefetch, LDGSTS = cp.async (generic-proxy staging of table rows, round 2), SYNCS = mbarrier operations, STAS = st.async to a peer CTA's shared
memory, DMMA = fp64 tensor-core MMA, plus the fp64 pipe, memory and barrier instructions.  No GPU needed.

    python scripts/sass_summary.py > profiles/sass_summary_r2.md
"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "ciaoalgorithms.jl_b200", "libciao_cuda.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
keys = ["UBLKCP", "UBLKPF", "LDGSTS", "SYNCS", "STAS", "DMMA", "DFMA", "DMUL", "DADD", "MUFU", "LDG", "STG", "LDS", "STS", "SHFL", "BAR.SYNC", "MEMBAR", "ATOM"]
want = {
    "row_pass_kernel<16, 0, 0>": "K1 full gradient, d = 4096, least squares (the roofline kernel)",
    "row_pass_kernel<4, 1, 1>": "K2 SAGA table init, d = 1024, logistic",
    "seq_kernel<4, 1, 0, 1, true>": "K3 SVRG inner epoch, d = 4096, LS + NormL1, cached c_i(z_full)",
    "seq_kernel<2, 2, 1, 1, false>": "K4 SAGA steps, d = 1024, logistic + NormL1",
    "seq_kernel<2, 3, 1, 1, false>": "K5 Finito steps, d = 1024, logistic + NormL1",
    "adaptive_kernel<2, 1, 1>": "adaptive Finito (on-device linesearch), logistic + NormL1",
    "batch_sm_kernel<8, 0, 1>": "Finito minibatches, one CTA per SM, flagged-word exchange (round 2), d = 1024, logistic",
    "batch_sm_kernel<8, 2, 1>": "LFinito minibatch sweep with cached c_i(z_full), same kernel",
    "batch_persistent_kernel<4, 0, 1>": "round-1 persistent minibatch kernel (grid barriers), kept as CIAO_BATCH_EXCHANGE=barrier",
    "proshi_steps_kernel<1, 2>": "K7 ProShI, batch 1, IndBox",
    "proshi_batch_kernel<512, 4>": "K7 ProShI, batches >= 64 blocks",
    "pass_tail_kernel": "tail of a pass: CTA-partial reduction + peer-memory exchange + closing update (round 2)",
    "exchange_floor_kernel<1>": "exchange-latency measurement (round 2)",
}
rows = {}
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    mangled = f.split("\n", 1)[0].strip()
    short = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "")
    if short in want and short not in rows:
        rows[short] = (len(re.findall(r"/\*[0-9a-f]{4}\*/", f)), {k: len(re.findall(r"\b" + re.escape(k) + r"[\.\s;]", f)) for k in keys})
print("# SASS inventory of the hot kernels (`scripts/sass_summary.py`, `cuobjdump -sass` of the in-tree libciao_cuda.so, sm_100a)\n")
print("Static instruction counts of one representative instantiation per kernel family.  UBLKCP = TMA bulk copy (`cp.async.bulk`), UBLKPF = TMA L2")
print("prefetch, LDGSTS = `cp.async` (the generic-proxy staging of table rows written inside the kernel, round 2), SYNCS = mbarrier operations, STAS = `st.async` into a peer CTA's shared memory (the DSMEM exchange), DMMA = fp64")
print("tensor-core MMA (the shuffle-free warp sums); SHFL = 0 in the sequential kernels — no shuffle is left on the step's path;")
print("ATOM = 0 everywhere — no floating-point atomics (fixed-order reductions, bitwise reproducible).\n")
print("| kernel | role | instr | " + " | ".join(keys) + " |")
print("|---|---|---|" + "---|" * len(keys))
for short, role in want.items():
    if short in rows:
        n, c = rows[short]
        print(f"| `{short}` | {role} | {n} | " + " | ".join(str(c[k]) for k in keys) + " |")
    else:
        print(f"| `{short}` | {role} | not found | " + " | ".join("" for _ in keys) + " |")

# where the fences of the SVRG kernel sit: none inside the step loop
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    if f.split("\n", 1)[0].strip() == "_Z10seq_kernelILi4ELi1ELi0ELi1ELb1EEv7SeqArgs":
        lines = [l for l in f.splitlines() if re.search(r"/\*[0-9a-f]{4}\*/", l)]
        addr = lambda l: int(re.search(r"/\*([0-9a-f]{4})\*/", l).group(1), 16)
        at = lambda key: [addr(l) for l in lines if re.search(r"\b" + key, l)]
        stas, membar, dmma = at("STAS"), at("MEMBAR"), at("DMMA")
        loops = []
        for l in lines:
            m = re.search(r"\bBRA\S*\s+.*?(0x[0-9a-f]+)", l)
            if m and int(m.group(1), 16) < addr(l):
                loops.append((int(m.group(1), 16), addr(l)))
        body = [lp for lp in loops if any(lp[0] <= a <= lp[1] for a in stas) and lp[1] - lp[0] < 0x3000]
        if body:
            lo, hi = min(body, key=lambda lp: lp[1] - lp[0])
            inside = [hex(a) for a in membar if lo <= a <= hi]
            print(f"\n`seq_kernel<4, 1, 0, 1, true>`: the step loop (two ping-pong steps per trip) spans {hex(lo)}–{hex(hi)} and holds "
                  f"{sum(lo <= a <= hi for a in dmma)} DMMA and {sum(lo <= a <= hi for a in stas)} STAS; MEMBAR inside the loop: {inside or 'none'} "
                  f"(all {len(membar)} MEMBARs sit in the prologue / epilogue cluster synchronisation: {', '.join(hex(a) for a in membar)}).")
        break
