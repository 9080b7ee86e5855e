"""Turn the ncu captures brought back in gpurun_out/ into the text summaries under profiles/.

    python tools/summarise_profiles.py <tag>       # e.g. final3 -> gpurun_out/prof_{c2,c4,c5}_<tag>.ncu-rep

Writes profiles/ncu_<workload>_r01_final.txt, profiles/traffic.json (DRAM bytes per launch),
profiles/launches_r01.csv and profiles/launch_share_r01.txt (from gpurun_out/launches_final.csv).
Needs the `ncu` CLI (reads reports; no GPU)."""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
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
DESC = {'c2': 'C2 (N=256, Pp=82; 128 threads x 4 CTAs/SM)', 'c4': 'C4 (N=512, Pp=162, coherent; 256 x 2)',
        'c5': 'C5 (N=1024, Pp=162; 256 x 3), 1 launch = 10000 pairs'}
UNIT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def raw_row(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    row = [r for r in rows[2:] if 'screen_detect' in ' '.join(r[:12])][0]
    return hdr, {h: (u, v) for h, u, v in zip(hdr, units, row)}


def main(tag):
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    traffic = json.load(open(tpath))
    for w in ('c2', 'c4', 'c5'):
        rep = os.path.join(ROOT, 'gpurun_out', f'prof_{w}_{tag}.ncu-rep')
        if not os.path.exists(rep):
            continue
        hdr, d = raw_row(rep)
        out = [f'# {DESC[w]}, FINAL round-1 kernel: packed-FP32 FFT, staged 2-row stores (N <= 512), window-specialised instance',
               f'# source: {os.path.basename(rep)} (ncu --set full --clock-control none --import-source on; one launch = one bench step)',
               '', 'kernel: ' + d['Kernel Name'][1]]
        out += [f"{k:80s} {d[k][1]:>20s} {d[k][0]}" for k in KEYS if k in d]
        out += ['', 'warp stall reasons (average warps stalled per issued instruction):']
        pre, suf = 'smsp__average_warps_issue_stalled_', '_per_issue_active.ratio'
        st = sorted(((float(d[h][1]), h) for h in hdr if h.startswith(pre) and h.endswith(suf) and 'not_issued' not in h),
                    reverse=True)
        out += [f"  {h[len(pre):-len(suf)]:28s} {v:.3f}" for v, h in st[:12]]
        open(os.path.join(ROOT, 'profiles', f'ncu_{w}_r01_final.txt'), 'w').write('\n'.join(out) + '\n')
        traffic[w] = sum(float(d[k][1]) * UNIT[d[k][0]] for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
        print(w, d['gpu__time_duration.sum'][1], 'ms; instr', d['smsp__inst_executed.sum'][1], 'ipc',
              d['sm__inst_executed.avg.per_cycle_elapsed'][1], 'data pipe %',
              d['l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'][1], 'DRAM GB', traffic[w] / 1e9)
    traffic['_comment'] = ('dram__bytes_read.sum + dram__bytes_write.sum per launch of the K2 kernel (bytes), from the ncu '
                           '--set full captures summarised in profiles/ncu_*_r01_final.txt; one launch = one bench step')
    json.dump(traffic, open(tpath, 'w'), indent=1)

    lpath = os.path.join(ROOT, 'gpurun_out', 'launches_final.csv')
    if os.path.exists(lpath):
        rows = [r for r in csv.reader(open(lpath)) if len(r) > 10]
        h = rows[0]
        ik, iv = h.index('Kernel Name'), h.index('Metric Value')
        tot = collections.OrderedDict()
        for r in rows[1:]:
            name = r[ik].replace('void ', '').replace('fastb::<unnamed>::', '').replace('<unnamed>::', '')
            short = name.split('(')[0][:78]
            tot.setdefault(short, [0, 0.0])
            tot[short][0] += 1
            tot[short][1] += float(r[iv].replace(',', '')) / 1e6
        allms = sum(v for _, v in tot.values())
        out = ['# ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 2 --warmup 3 --no-cpu --no-comparator',
               '# (final round-1 kernel; init + warm-up + timed steps; profiles/launches_r01.csv is the raw list)',
               '# per-launch times are cold-cache and serialised: compare SHARES, not absolutes', '']
        out += [f"{k:78s} launches {n:3d}  total {ms:9.3f} ms  share {100 * ms / allms:5.1f}%"
                for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1])]
        open(os.path.join(ROOT, 'profiles', 'launch_share_r01.txt'), 'w').write('\n'.join(out) + '\n')
        open(os.path.join(ROOT, 'profiles', 'launches_r01.csv'), 'w').write(open(lpath).read())


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'final')
