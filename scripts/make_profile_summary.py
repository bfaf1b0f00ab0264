"""Build profiles/ncu_<tag>_summary.{md,json} from gpurun_out/ artefacts (ncu raw CSV + launch list)."""
import csv, json, sys
tag = sys.argv[1] if len(sys.argv) > 1 else 'r1'
rows = list(csv.reader(open(f'gpurun_out/prof_{tag}_raw.csv')))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
want = [('gpu__time_duration.sum', 'duration_us'), ('launch__registers_per_thread', 'regs'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps_active_pct'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_throughput_pct'),
        ('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'fma_pipe_pct'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_throughput_pct'),
        ('dram__bytes_read.sum', 'dram_read_MB'), ('dram__bytes_write.sum', 'dram_write_MB'),
        ('smsp__inst_executed.sum', 'warp_instructions'), ('launch__waves_per_multiprocessor', 'waves')]
def table(rows):
    hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        d = {'kernel': r[idx['Kernel Name']]}
        for m, n in want:
            d[n] = r[idx[m]] if m in idx else None
        out.append(d)
    return out
kern = table(rows)
import os
mr_path = f'gpurun_out/prof_mr_{tag}_raw.csv'
mr = table(list(csv.reader(open(mr_path)))) if os.path.exists(mr_path) else []
lr = [r for r in csv.reader(open(f'gpurun_out/launches_step_{tag}.csv')) if len(r) > 5]
h2 = lr[0]; ki = h2.index('Kernel Name'); vi = h2.index('Metric Value')
names = [(r[ki], float(r[vi].replace(',', ''))) for r in lr[1:]]
st = [i for i, (n, _) in enumerate(names) if 'prepare2_kernel' in n][-1]   # first kernel of the last step
step = names[st:]
tot = sum(v for _, v in step)
md = [f'# profiles/ - round {tag} (B200, ncu, `--clock-control none`)\n',
      'Workload: BASELINE configs[1] - AdvancedMixConsole fwd+bwd + MRSTFT, B=8, N=16, T=262144, float32, bus-only mode.\n',
      'Commands (under gpurun):\n```\n'
      f'ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step_{tag}.csv python scripts/prof_step.py 3\n'
      f'ncu --set full --clock-control none --import-source on -k regex:"console_fwd|chain_bwd|track_bwd2" -s 3 -c 3 -o gpurun_out/prof_{tag} python scripts/prof_once.py 8 16 262144 2\n'
      f'ncu --set full --clock-control none --import-source on -k regex:"stft_loss|istft_grad|ola_multi" -s 8 -c 8 -o gpurun_out/prof_mr_{tag} python scripts/prof_mrstft.py 2\n'
      'python scripts/timeline_step.py   # torch.profiler / CUPTI timeline of the graph-replayed step: what overlaps\n'
      'python scripts/sweep_eq_comp.py ; python tests/tools/conv_bench.py\npython bench.py --steps 50 --warmup 5 ; python bench.py --impl reference --steps 3 --warmup 1\n```\n',
      '## 1. Launch list of one step (ncu per-launch times are cold-cache and SERIALISED: under ncu the three MRSTFT resolutions run one after the other and the master-bus backward kernel, which in the step runs on 4 B = 32 SMs beside the track kernel, is timed alone on those 32 SMs; `timeline_step_r2.txt` has the real overlap)\n',
      '| us | share | kernel |\n|---:|---:|---|']
for n, v in step:
    md.append(f'| {v/1000:.1f} | {100*v/tot:.1f}% | `{n[:110]}` |')
md.append(f'| **{tot/1000:.1f}** | 100% | total (device-timed step without a profiler: `bench_{tag}.json` `ms_per_step`) |\n')
md.append('## 2. `ncu --set full` of the three console chain kernels (one launch each)\n')
md.append('| kernel | us | regs | warps active % | SM throughput % | FMA pipe % | DRAM throughput % | DRAM read MB | DRAM write MB | warp-instructions |\n|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|')
for d in kern:
    f = lambda k: float(str(d[k]).replace(',', ''))
    md.append(f"| `{d['kernel'][:60]}` | {f('duration_us'):.1f} | {d['regs']} | {f('warps_active_pct'):.1f} | {f('sm_throughput_pct'):.1f} | {f('fma_pipe_pct'):.1f} | {f('dram_throughput_pct'):.1f} | {f('dram_read_MB'):.1f} | {f('dram_write_MB'):.1f} | {f('warp_instructions'):.0f} |")
def rowfmt(d):
    f = lambda k: float(str(d[k]).replace(',', ''))
    return (f"| `{d['kernel'][:60]}` | {f('duration_us'):.1f} | {d['regs']} | {f('warps_active_pct'):.1f} | {f('sm_throughput_pct'):.1f} | "
            f"{f('fma_pipe_pct'):.1f} | {f('dram_throughput_pct'):.1f} | {f('dram_read_MB'):.1f} | {f('dram_write_MB'):.1f} | {f('warp_instructions'):.0f} |")
if mr:
    md.append('\n## 3. `ncu --set full` of the MRSTFT kernels (one resolution each; run alone, cold cache)\n')
    md.append('| kernel | us | regs | warps active % | SM throughput % | FMA pipe % | DRAM throughput % | DRAM read MB | DRAM write MB | warp-instructions |\n|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|')
    for d in mr:
        md.append(rowfmt(d))
open(f'profiles/ncu_{tag}_summary.md', 'w').write('\n'.join(md) + '\n')
json.dump({'kernels': kern, 'mrstft_kernels': mr, 'step_launches_us': [(n, v / 1000) for n, v in step]}, open(f'profiles/ncu_{tag}_summary.json', 'w'), indent=1)
print('\n'.join(md[-7:]))
