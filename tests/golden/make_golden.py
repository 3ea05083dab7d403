"""Generate golden fixtures by EXECUTING THE REFERENCE'S OWN PYTHON in this container.

The reference package cannot be imported (flashfftconv/conv.py:9 imports the compiled `monarch_cuda`
extension and there is no GPU here), so this script slices the pure-PyTorch functions out of the
reference sources with `ast` and executes them unmodified:

  * `ref_fft_conv`                      /root/reference/tests/test_flashfftconv.py:5-13  (the test oracle)
  * `fft_matrix`, `ifft_matrix`,
    `compute_twiddle_factors_fft/ifft`  /root/reference/flashfftconv/conv.py:22-52       (constant tables)

Outputs (committed): tests/golden/conv_*.npz (inputs + reference outputs, fwd and autograd grads) and
tests/golden/tables.npz.  /root/reference does not exist on the GPU box, so tests only read the .npz.
Run:  python tests/golden/make_golden.py
"""
import ast
import os

import numpy as np
import torch

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))


def slice_functions(path, names):
    src = open(path).read()
    tree = ast.parse(src)
    ns = {'torch': torch, 'math': __import__('math')}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            code = compile(ast.Module(body=[node], type_ignores=[]), path, 'exec')
            exec(code, ns)
    missing = [n for n in names if n not in ns]
    assert not missing, missing
    return ns


def main():
    t = slice_functions(os.path.join(REF, 'tests/test_flashfftconv.py'), ['ref_fft_conv'])
    c = slice_functions(os.path.join(REF, 'flashfftconv/conv.py'),
                        ['fft_matrix', 'ifft_matrix', 'compute_twiddle_factors_fft', 'compute_twiddle_factors_ifft'])
    ref_fft_conv = t['ref_fft_conv']

    # ---- tables (conv.py:132-156 uses exactly these calls for seqlen 8192)
    np.savez_compressed(
        os.path.join(HERE, 'tables.npz'),
        f_32=c['fft_matrix'](32).numpy(), f_16=c['fft_matrix'](16).numpy(),
        if_32=c['ifft_matrix'](32).numpy(), if_16=c['ifft_matrix'](16).numpy(),
        tw_16_16=c['compute_twiddle_factors_fft'](16, 16).numpy(),
        tw_32_256=c['compute_twiddle_factors_fft'](32, 256).numpy(),
        itw_16_16=c['compute_twiddle_factors_ifft'](16, 16).numpy(),
        itw_32_256=c['compute_twiddle_factors_ifft'](32, 256).numpy())

    # ---- convolution cases: (name, B, H, N, L, dtype, gated), inputs drawn like the reference tests
    cases = [
        ('n1024_fp32', 2, 16, 1024, 1024, torch.float32, False),      # BASELINE config 1
        ('n256_bf16', 2, 3, 256, 256, torch.bfloat16, False),
        ('n4096_bf16_pad', 2, 3, 4096, 2048, torch.bfloat16, False),
        ('n8192_bf16', 3, 4, 8192, 8192, torch.bfloat16, False),       # odd batch
        ('n8192_bf16_pad', 2, 4, 8192, 4096, torch.bfloat16, False),
        ('n8192_bf16_gated', 2, 4, 8192, 4096, torch.bfloat16, True),
        ('n8192_fp16', 2, 2, 8192, 8192, torch.float16, False),
        ('n32768_bf16_gated_pad', 2, 2, 32768, 16384, torch.bfloat16, True),   # shape of config 3
    ]
    for name, B, H, N, L, dtype, gated in cases:
        torch.manual_seed(0)                                           # tests/test_flashfftconv.py:54
        u = (torch.randn(B, H, L) * 0.02).to(dtype)
        k = torch.randn(H, L) * 0.02 * torch.exp(-0.1 * torch.arange(L))
        dout = (torch.randn(B, H, L) * 0.02).to(dtype)
        u_ = u.clone().requires_grad_(True)
        k_ = k.clone().requires_grad_(True)
        save = {}
        if gated:
            pre = (torch.randn(B, H, L) * 0.02).to(dtype).requires_grad_(True)
            post = (torch.randn(B, H, L) * 0.02).to(dtype).requires_grad_(True)
            y = ref_fft_conv(u_ * pre, k_, N) * post                   # tests/test_flashfftconv.py:208
        else:
            y = ref_fft_conv(u_, k_, N)
        y.backward(dout)                                               # tests/test_flashfftconv.py:100
        save.update(u=u.float().numpy(), k=k.numpy(), dout=dout.float().numpy(), y=y.detach().float().numpy(),
                    du=u_.grad.float().numpy(), dk=k_.grad.numpy(), N=np.int64(N),
                    dtype=np.array(str(dtype)))
        if gated:
            save.update(pregate=pre.detach().float().numpy(), postgate=post.detach().float().numpy(),
                        dpregate=pre.grad.float().numpy(), dpostgate=post.grad.float().numpy())
        np.savez_compressed(os.path.join(HERE, f'conv_{name}.npz'), **save)
        print(name, 'max|y|', float(y.abs().max()))


if __name__ == '__main__':
    main()
