#!/usr/bin/env python
"""Turn the two ncu captures of a round into the committed summaries under profiles/:
  tools/summarize_ncu.py <launches.csv> <full.ncu-rep> <tag>
writes profiles/<tag>_ncu_launches.csv (copy), <tag>_ncu_launch_shares.csv, <tag>_ncu_full_top_kernels.json and refreshes
profiles/ncu_traffic.json (DRAM bytes per launch per kernel family, read by bench.py for roofline.traffic)."""
import collections, csv, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
launches, rep, tag = sys.argv[1:4]
prof = os.path.join(ROOT, 'profiles')
shutil.copy(launches, os.path.join(prof, f'{tag}_ncu_launches.csv'))
rows = list(csv.reader(open(launches)))
hdr = [r for r in rows if r and r[0] == 'ID'][0]
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
body = [r for r in rows if r and r[0].isdigit()]
for r in body:
    k = r[ix['Kernel Name']].split('(')[0]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += float(r[ix['Metric Value']]) / 1e3
tot = sum(v[1] for v in agg.values())
with open(os.path.join(prof, f'{tag}_ncu_launch_shares.csv'), 'w') as fp:
    fp.write('# ncu launch list (gpu__time_duration.sum, --clock-control none) of bench.py: per-kernel share of %d launches\n' % len(body))
    fp.write('# cold-cache, serialised launches: compare SHARES with the CUDA-event kernel_families of the bench line, not absolutes\n')
    fp.write('kernel,launches,total_us,share\n')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        fp.write('%s,%d,%.1f,%.4f\n' % (k, v[0], v[1], v[1] / tot))
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h2 = rr[0]
want = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum']
idx = [h2.index(w) for w in want if w in h2]
full = [{h2[i]: r[i] for i in idx} for r in rr[2:]]
json.dump(full, open(os.path.join(prof, f'{tag}_ncu_full_top_kernels.json'), 'w'), indent=1)
fam = {'k_attn_tc': ['attention_tc'], 'k_gemm_tc<4>': ['gemm_qkvx_pos'], 'k_gemm_tc<1>': ['gemm_ffn_up'], 'k_gemm_tc<6>': ['gemm_pw1_glu'],
       'k_dw_fast': ['dwconv_bn_silu_bf16'], 'k_dw_v2': ['dwconv_bn_silu_bf16'], 'k_frontend_convs2': ['frontend_convs'], 'k_frontend_tc': ['frontend_tc'], 'k_ffn_fused': ['ffn_fused', 'ffn_fused_final']}
src = f'profiles/{tag}_ncu_full_top_kernels.json (ncu --set full, B=256 W=1536)'
def dram(o):
    return (float(o['dram__bytes_read.sum']) + float(o['dram__bytes_write.sum'])) * 1e6
traffic = {}
for o in full:
    for k, fs in fam.items():
        if k in o['Kernel Name']:
            for f in fs:
                traffic[f] = dict(dram_bytes_per_launch=dram(o), ncu_duration_us=float(o['gpu__time_duration.sum']), source=src)
res = [o for o in full if 'k_gemm_tc<5>' in o['Kernel Name']]
big = [o for o in res if float(o['dram__bytes_read.sum']) > 120]
small = [o for o in res if float(o['dram__bytes_read.sum']) <= 120]
for names, os_ in ((('gemm_ffn_down', 'gemm_ffn_down_final'), big), (('gemm_attn_out', 'gemm_pw2'), small)):
    if os_:
        for n in names:
            traffic[n] = dict(dram_bytes_per_launch=sum(dram(o) for o in os_) / len(os_),
                              ncu_duration_us=sum(float(o['gpu__time_duration.sum']) for o in os_) / len(os_), source=src)
json.dump(traffic, open(os.path.join(prof, 'ncu_traffic.json'), 'w'), indent=1)
for o in full:
    print(o['Kernel Name'][:44], o['gpu__time_duration.sum'], 'us  dram rd/wr MB', o['dram__bytes_read.sum'][:6], o['dram__bytes_write.sum'][:6],
          ' dram%', o['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'][:5], ' tensor%', o['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'][:5],
          ' issue%', o['smsp__issue_active.avg.pct_of_peak_sustained_active'][:5], ' xu%', o['sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'][:5])
