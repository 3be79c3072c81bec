#!/usr/bin/env python3
"""Register-file read traffic of an ensemble kernel from an ncu report (source page, SASS with execution counts).

Model (tools/microbench/fp64_operands.cu, measured on B200): a scheduler's register file delivers two 32-bit vector-register
source operands per lane and cycle; a DADD/DMUL with two register sources needs 2 cycles (= its FP64 pipe time), a DFMA with
three register sources 3 cycles, and every register source of the ALU/MOV/select instructions around them adds to the same
budget.  Uniform-register, constant and immediate operands are free, and so are operands served by the reuse cache (.reuse).
Prints the predicted cycles per scheduler against the measured ones.   Usage: tools/sass_regread.py report.ncu-rep"""
import collections, csv, re, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
rawm = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
mr = list(csv.reader(rawm.splitlines()))
met = dict(zip(mr[0], mr[2]))
cycles = float(met['smsp__cycles_active.avg'].replace(',', '')) if 'smsp__cycles_active.avg' in met else float(met['sm__cycles_active.avg'].replace(',', ''))
n_smsp = 148 * 4
WIDE = ('DADD', 'DMUL', 'DFMA', 'DSETP', 'MUFU.RCP64H', 'MUFU.RSQ64H')
tot_words = 0.0; tot_inst = 0.0; fp64_inst = 0.0; fp64_pipe = 0.0
by_op = collections.Counter(); cnt_op = collections.Counter()
reuse_next = {}
for r in data:
    src = r[ix['Source']].strip()
    try: n = float(r[ix['Instructions Executed']])
    except ValueError: continue
    if n == 0: continue
    toks = src.replace(',', ' ').replace(';', ' ').split()
    if toks and toks[0].startswith('@'): toks = toks[1:]
    if not toks: continue
    op = toks[0]
    base = op.split('.')[0]
    wide = op.startswith('D') and base in ('DADD', 'DMUL', 'DFMA', 'DSETP')
    ops = toks[1:]
    # destination(s): first operand for most instructions; predicates for *SETP (2 dests); stores/branches have none
    if base in ('STS', 'STG', 'STL', 'ST', 'BRA', 'BSSY', 'BSYNC', 'EXIT', 'RED', 'ATOMS', 'WARPSYNC', 'NOP', 'BAR', 'MEMBAR', 'ERRBAR', 'CCTL'):
        srcs = ops
    elif base.endswith('SETP') or base in ('PLOP3',):
        srcs = ops[2:]
    else:
        srcs = ops[1:]
    words = 0
    for o in srcs:
        m = re.search(r'(?<![UP])R(\d+)', o)   # vector register (not UR / PR)
        if not m or o.startswith('UR') or 'c[' in o: continue
        if m.group(1) == 'Z': continue
        if '.reuse' in o: continue
        w = 2 if (wide or '.64' in op or base in ('DADD', 'DMUL', 'DFMA', 'DSETP')) else 1
        if base == 'MUFU': w = 1
        if base in ('STG', 'STS', 'LDG', 'LDS', 'STL', 'LDL') and '[' in o: w = 1 if base in ('STS', 'LDS', 'STL', 'LDL') else 2
        words += w
    tot_words += words * n; tot_inst += n
    by_op[base] += words * n; cnt_op[base] += n
    if base in ('DADD', 'DMUL', 'DFMA', 'DSETP'):
        fp64_inst += n
print("warp instructions %.4g, FP64 %.4g (%.1f%%)" % (tot_inst, fp64_inst, fp64_inst / tot_inst * 100))
print("register source words (32-bit, per lane) per warp instruction: %.3f" % (tot_words / tot_inst))
pred_rf = tot_words / 2 / n_smsp
pred_pipe = fp64_inst * 2 / n_smsp
print("cycles per scheduler: measured %.4g | register-file model (words/2) %.4g = %.1f%% | FP64 pipe (2 per instruction) %.4g = %.1f%% | issue (1 per instruction) %.4g = %.1f%%"
      % (cycles, pred_rf, pred_rf / cycles * 100, pred_pipe, pred_pipe / cycles * 100, tot_inst / n_smsp, tot_inst / n_smsp / cycles * 100))
print("register words by opcode (share of all words | words per instruction):")
for k, v in by_op.most_common(14):
    print("  %-8s %5.1f%%  %.2f" % (k, v / tot_words * 100, v / cnt_op[k]))
