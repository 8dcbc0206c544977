"""Run each GPU self-test in its own process (a trapping kernel must not poison the others) and
write a JSON summary to gpurun_out/probe.json. Usage: python tools/gpu_probe.py"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [(64, 256, 0, 1), (64, 256, 0, 3), (128, 256, 0, 3), (128, 16, 0, 3), (64, 64, 0, 3),
         (64, 256, 1, 1), (64, 256, 1, 3), (128, 256, 1, 3), (128, 16, 1, 3), (64, 64, 1, 3),
         (64, 256, 2, 1), (64, 256, 2, 3), (128, 256, 2, 3), (128, 16, 2, 3), (64, 144, 2, 3),
         (64, 256, 3, 3), (128, 256, 3, 3), (128, 16, 3, 3), (64, 144, 3, 3)]

CODE = """
import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)
from test_tc_selftest_gpu import run_selftest
print('ERR=%%.3e' %% run_selftest(%d, %d, %d, %d))
"""


def main():
    out = []
    for (K, N, mode, passes) in CASES:
        code = CODE % (ROOT, os.path.join(ROOT, 'tests'), K, N, mode, passes)
        try:
            p = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=120)
            msg = (p.stdout + p.stderr).strip().splitlines()
            err = [l for l in msg if l.startswith('ERR=')]
            out.append({'K': K, 'N': N, 'mode': mode, 'passes': passes, 'rc': p.returncode,
                        'err': err[0] if err else None, 'tail': msg[-3:] if not err else []})
        except subprocess.TimeoutExpired:
            out.append({'K': K, 'N': N, 'mode': mode, 'passes': passes, 'rc': 'timeout'})
        print(out[-1], flush=True)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'probe.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
