"""profiles/r2_ref_gpu.json (written by baseline/run_ref.py on the GPU box) -> profiles/r2_ref_gpu.md."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = json.load(open(os.path.join(ROOT, 'profiles', 'r2_ref_gpu.json')))
out = []
w = out.append
w('# r2 — the reference\'s own CUDA kernels on the same B200 (baseline/build_ref.py, baseline/run_ref.py)\n')
w('The unmodified reference extension (`csrc/flashfftconv`, 25 sources, `-O3 --use_fast_math -std=c++17`, torch\'s CUDAExtension')
w('defines) compiled here to sm_100 SASS (wmma `HMMA` code; the reference\'s own setup.py emits compute_80 PTX only) and run on the')
w('GPU box next to ours on identical inputs.  Raw: `profiles/r2_ref_gpu.json`; this file: `python baseline/ref_gpu_md.py`.\n')
w('## Output agreement (rel-L2; bf16; inputs u ~ N(0,1), k ~ N(0,1/L))\n')
w('| N | B | H | L | gated | ours vs fp32 torch.fft | reference vs fp32 | ours vs reference (y / du / dk [/ dpregate / dpostgate]) |')
w('|---|---|---|---|---|---|---|---|')
for a in d['agreement']:
    o = a['ours_vs_ref']
    vals = ' / '.join('%.2e' % o[k] for k in ('y', 'du', 'dk', 'dpregate', 'dpostgate') if k in o)
    w(f"| {a['N']} | {a['B']} | {a['H']} | {a['L']} | {a['gated']} | {a['ours_vs_fp32']:.2e} | {a['ref_vs_fp32']:.2e} | {vals} |")
w('\nOurs is closer to the fp32 truth than the reference everywhere (fp32 twiddle / k_f multiplies vs the reference\'s bf16 ones).')
w('The reference\'s gated N=32768, L=N/2 `du` differs from ours (and from autograd through the fp32 oracle, which ours matches to')
w('5.6e-3 at the full C3 shape, `profiles/parity_r2.md`) by O(1): its other four outputs of the same call agree.\n')
w('## Timing, same box, module call incl. k -> k_f, training mode (ms; CUDA events, 20 calls after warm-up)\n')
w('| config | shape | ours fwd | reference fwd | x | ours fwd+bwd | reference fwd+bwd | x |')
w('|---|---|---|---|---|---|---|---|')
for t in d['timing']:
    shape = f"N={t['N']} B={t['B']} H={t['H']} L={t['L']}{' gated' if t['gated'] else ''}"
    w(f"| {t['config']} | {shape} | {t['ours_fwd_ms']:.3f} | {t['ref_fwd_ms']:.3f} | {t['ref_fwd_ms'] / t['ours_fwd_ms']:.1f} | "
      f"{t['ours_fwd_bwd_ms']:.3f} | {t['ref_fwd_bwd_ms']:.3f} | {t['ref_fwd_bwd_ms'] / t['ours_fwd_bwd_ms']:.1f} |")
w('\n(C5 at H=16: the reference needs H % 16 == 0 above 32K, README.md:269; `bench.py` times the H=8 per-GPU shard.  r256 … r8k:')
w('the shape of the reference\'s published table, README.md:224-231 — gated forward, B=64, H=768; its H100-SXM fp16 figures there:')
w('0.11 ms at 256, 0.29 ms at 1K, 1.43 ms at 4K, 3.58 ms at 8K.)')
open(os.path.join(ROOT, 'profiles', 'r2_ref_gpu.md'), 'w').write('\n'.join(out) + '\n')
print('\n'.join(out[-16:]))
