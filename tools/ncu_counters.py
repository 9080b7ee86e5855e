"""ncu reports (gpurun_out/prof_<workload>_<tag>.ncu-rep) -> the committed evidence of a round:

    python tools/ncu_counters.py <tag> [round]      # e.g. r02a 02

  profiles/ncu_<workload>_r<round>.txt        raw counters of the K2 launch (time, DRAM, L2, pipes, stalls)
  profiles/sass_mix_<workload>_r<round>.txt   dynamic SASS instruction mix per pair: which part of the
                                              kernel (Philox / Box-Muller / FFT / memory / detector) executes
                                              which opcodes, from the source-correlated page (-lineinfo)
  profiles/kernel_counters_r<round>.json      per-pair DRAM bytes and warp instructions, keyed by the
                                              digest of the kernel sources they were captured from; bench.py
                                              uses them for roofline.traffic / roofline.secondary only while
                                              the digest still matches the tree

Needs the `ncu` CLI (reads reports; no GPU)."""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

KEYS = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'launch__grid_size', 'launch__block_size',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors.sum', 'lts__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum']
UNIT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
# pairs per captured launch: the timed step of `bench.py --workload <w>` (WORKLOADS in bench.py)
PAIRS = {'c2': 600000, 'c4': 150000, 'c5': 40000, 'c2fast': 600000, 'c1prime': 100000}


def ncu_csv(rep, *args):
    out = subprocess.run(['ncu', '-i', rep, *args, '--csv'], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def raw_counters(rep):
    rows = ncu_csv(rep, '--page', 'raw')
    hdr, units = rows[0], rows[1]
    row = [r for r in rows[2:] if 'screen_detect' in ' '.join(r[:12])][-1]
    return {h: (u, v) for h, u, v in zip(hdr, units, row)}


def num(d, key):
    u, v = d[key]
    return float(v.replace(',', '')) * UNIT.get(u, 1)


# ---- dynamic instruction mix from the source-correlated page ---------------------------------------
def function_ranges(path):
    """[(first line, last line, name)] of the top-level functions / lambdas of one of our sources: a
    line that starts a definition (`... name(...) {` at brace depth <= 1) opens a range."""
    out, stack = [], []
    try:
        lines = open(path).read().split('\n')
    except OSError:
        return out
    pat = re.compile(r'\b([A-Za-z_][A-Za-z_0-9]*)\s*\([^;]*$')
    for i, l in enumerate(lines, 1):
        s = l.strip()
        if s.startswith(('FASTB_HD', '__device__', '__global__', 'template', 'static', 'inline')) or \
                re.match(r'^(void|int|float2?|uint4|pc)\b', s):
            m = pat.search(s)
            if m and not s.startswith('template'):
                if stack:
                    out.append((stack[0], i - 1, stack[1]))
                stack = [i, m.group(1)]
    if stack:
        out.append((stack[0], len(lines), stack[1]))
    return out


CATEGORY = [
    (r'philox|noise_block_fields', 'Philox + bit fields'),
    (r'weighted_normal|lg2_ftz|box_muller', 'Box-Muller'),
    (r'chi_normal', 'epilogue'),
    (r'dft|cmul|cadd|csub|add2|sub2|mul2|fma2|fnma2|neg2|bc2|phase_|run|apply_twiddle|get16|put|get|k_off|k_base|make_tw|LineFFT', 'FFT'),
    (r'accumulate|sh_phase|cmac', 'detector'),
    (r'finish_pair|stats_|warp_sum|pair_id|atomic_', 'epilogue'),
]


LINE_HINTS = [
    (r'noise_block_fields', 'Philox + bit fields'),
    (r'weighted_normal', 'Box-Muller'),
    (r'run_shfl|F::run|convolve\(', 'FFT'),
    (r'accumulate\(|sh_phase', 'detector'),
    (r'finish_pair|stats_flush|sh_prepare', 'epilogue'),
    (r'__stcg|__ldg|__ldcg|tile\[|tl\[|rows\[', 'loads / stores (weights, scratch, pupil)'),
]


def categorise(fname, func, opcode, text=''):
    base = os.path.basename(fname)
    if base == 'fft_core.cuh':
        return 'FFT'
    if func and func.startswith('screen_detect_'):
        # a line of the kernel itself: the call site names the part when the callee's own lines do not
        for pat, cat in LINE_HINTS:
            if re.search(pat, text):
                return cat
    for pat, cat in CATEGORY:
        if re.search(pat, func or ''):
            if cat == 'FFT' and base != 'fft_core.cuh':
                continue
            return cat
    if re.match(r'(LD|ST|ATOM|RED|LDG|STG|LDS|STS)', opcode):
        return 'loads / stores (weights, scratch, pupil)'
    return 'kernel body (indexing, control, crop masks)'


def opclass(op):
    op = op.split('.')[0]
    if op in ('IMAD', 'IADD3', 'IADD', 'LEA', 'SHF', 'LOP3', 'PRMT', 'ISETP', 'SEL', 'IABS', 'SGXT', 'VIADD', 'VIMNMX',
              'UIADD3', 'ULEA', 'UMOV', 'USHF', 'ULOP3', 'UIMAD', 'UISETP', 'USEL', 'IMNMX', 'POPC', 'FLO', 'BREV', 'LOP'):
        return 'integer'
    if op in ('FADD', 'FMUL', 'FFMA', 'FADD2', 'FMUL2', 'FFMA2', 'FSEL', 'FSETP', 'FMNMX', 'FCHK'):
        return 'fp32 ' + ('packed' if op.endswith('2') else 'scalar')
    if op == 'MUFU':
        return 'MUFU'
    if op in ('LDG', 'STG', 'LD', 'ST', 'ATOM', 'ATOMG', 'RED', 'LDC', 'LDCU', 'ULDC'):
        return 'global / const memory'
    if op in ('LDS', 'STS', 'LDSM'):
        return 'shared memory'
    if op in ('SHFL',):
        return 'shuffle'
    if op in ('BAR', 'BRA', 'BSSY', 'BSYNC', 'EXIT', 'WARPSYNC', 'NOP', 'CALL', 'RET', 'BRX', 'DEPBAR', 'YIELD', 'BMOV'):
        return 'control / sync'
    if op in ('MOV', 'S2R', 'CS2R', 'S2UR', 'R2UR', 'I2F', 'F2I', 'I2FP', 'F2FP', 'F2F', 'R2P', 'P2R', 'PLOP3'):
        return 'moves / conversions'
    if op.startswith('D'):
        return 'fp64'
    return 'other'


PRIORITY = ['Philox + bit fields', 'Box-Muller', 'detector', 'epilogue', 'FFT',
            'loads / stores (weights, scratch, pupil)', 'kernel body (indexing, control, crop masks)']


def sass_mix(rep):
    """The correlated view lists an inlined instruction under every source line of its inline chain
    (callee line and each call site), so instructions are de-duplicated by address and attributed to
    the most specific part that claims them (PRIORITY order)."""
    rows = ncu_csv(rep, '--page', 'source', '--print-source', 'cuda,sass')
    seen = {}
    fname, ranges, cur_line, cur_text = None, [], 0, ''
    col = None
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            fname = r[1]
            local = os.path.join(ROOT, 'fast_b200', 'csrc', os.path.basename(fname))
            ranges = function_ranges(local)
            continue
        if r[0] == 'Line No':
            col = r.index('Instructions Executed')
            continue
        if r[0] == 'Function Name' or col is None or fname is None:
            continue
        if r[0] not in ('', '-'):
            try:
                cur_line = int(r[0])
                cur_text = r[1]
            except ValueError:
                pass
            continue
        sass = r[3].strip()
        if not sass or sass == '...':
            continue
        try:
            n = int(r[col])
        except (ValueError, IndexError):
            continue
        toks = sass.split()
        op = toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]
        func = next((nm for a, b, nm in ranges if a <= cur_line <= b), None)
        ours = os.path.exists(os.path.join(ROOT, 'fast_b200', 'csrc', os.path.basename(fname)))
        cat = categorise(fname, func, op, cur_text) if ours else None
        addr = int(r[2], 16)
        if addr not in seen or (cat is not None and (seen[addr][0] is None or
                                                     PRIORITY.index(cat) < PRIORITY.index(seen[addr][0]))):
            seen[addr] = (cat, op, n)
    # intrinsics (__fadd2_rn, __shfl_xor_sync, ... in the CUDA headers) carry only their own header line:
    # they inherit the part of the nearest instruction before them in address order that has one
    mix = collections.defaultdict(lambda: collections.Counter())
    last = 'kernel body (indexing, control, crop masks)'
    for addr in sorted(seen):
        cat, op, n = seen[addr]
        if cat is None:
            cat = last
        else:
            last = cat
        mix[cat][opclass(op)] += n
    return mix


def write_mix(path, title, mix, pairs):
    classes = sorted({c for m in mix.values() for c in m})
    total = sum(sum(m.values()) for m in mix.values())
    out = [title, '# warp-level instructions executed PER PAIR (one complex transform = two realisations), '
                  'from the ncu source page correlated through -lineinfo; rows: part of the kernel, columns: opcode class', '']
    out.append(f"{'part':52s} {'total':>10s} {'%':>6s}  " + ' '.join(f'{c[:14]:>14s}' for c in classes))
    for cat, m in sorted(mix.items(), key=lambda kv: -sum(kv[1].values())):
        t = sum(m.values())
        out.append(f"{cat:52s} {t / pairs:10.0f} {100 * t / total:6.1f}  " + ' '.join(f'{m.get(c, 0) / pairs:14.0f}' for c in classes))
    out.append(f"{'all':52s} {total / pairs:10.0f} {100.0:6.1f}  " +
               ' '.join(f'{sum(m.get(c, 0) for m in mix.values()) / pairs:14.0f}' for c in classes))
    open(path, 'w').write('\n'.join(out) + '\n')
    return total / pairs


def main(tag, rnd):
    import bench
    rec = {'_comment': 'per-pair counters of the K2 launch captured by ncu --set full (one launch = the timed step of '
                       'bench.py --workload <w>); used by bench.py only while kernel_digest matches the tree',
           'kernel_digests': {k: bench.kernel_digest(k) for k in ('radix', 'bluestein')}, 'workloads': {}}
    for w, pairs in PAIRS.items():
        rep = os.path.join(ROOT, 'gpurun_out', f'prof_{w}_{tag}.ncu-rep')
        if not os.path.exists(rep):
            continue
        d = raw_counters(rep)
        out = [f'# {w}: K2 launch of {pairs} pairs, round {rnd} kernel', f'# source: {os.path.basename(rep)} '
               '(ncu --set full --clock-control none --import-source on)', '', 'kernel: ' + d['Kernel Name'][1]]
        out += [f"{k:80s} {d[k][1]:>20s} {d[k][0]}" for k in KEYS if k in d]
        out += ['', 'warp stall reasons (average warps stalled per issued instruction):']
        stall = [(k, d[k][1]) for k in d if k.startswith('smsp__average_warp') and 'per_issue_active' in k and
                 'not_issued' not in k]
        out += [f"  {k:90s} {v}" for k, v in sorted(stall, key=lambda kv: -float(kv[1].replace(',', '') or 0))[:10]]
        open(os.path.join(ROOT, 'profiles', f'ncu_{w}_r{rnd}.txt'), 'w').write('\n'.join(out) + '\n')
        mix = sass_mix(rep)
        per_pair = write_mix(os.path.join(ROOT, 'profiles', f'sass_mix_{w}_r{rnd}.txt'),
                             f'# {w}: dynamic SASS instruction mix of {d["Kernel Name"][1][:120]}', mix, pairs)
        dram = num(d, 'dram__bytes_read.sum') + num(d, 'dram__bytes_write.sum')
        rec['workloads'][w] = {'pairs_per_launch': pairs, 'dram_bytes_per_pair': dram / pairs,
                               'dram_read_bytes_per_pair': num(d, 'dram__bytes_read.sum') / pairs,
                               'dram_write_bytes_per_pair': num(d, 'dram__bytes_write.sum') / pairs,
                               'warp_inst_per_pair': num(d, 'smsp__inst_executed.sum') / pairs,
                               'warp_inst_per_pair_source_page': per_pair,
                               'kernel_ms_under_ncu': num(d, 'gpu__time_duration.sum') * (1e-6 if d['gpu__time_duration.sum'][0] in ('ns', 'nsecond') else 1e-3 if d['gpu__time_duration.sum'][0] in ('us', 'usecond') else 1.0),
                               'l2_hit_pct': num(d, 'lts__t_sector_hit_rate.pct'),
                               'ipc': num(d, 'sm__inst_executed.avg.per_cycle_elapsed'),
                               'kernel': d['Kernel Name'][1]}
        print(w, json.dumps(rec['workloads'][w])[:300])
    json.dump(rec, open(os.path.join(ROOT, 'profiles', f'kernel_counters_r{rnd}.json'), 'w'), indent=1)


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else '02')
