"""Builds libbyolo.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), 'csrc')
LIB = os.path.join(HERE, 'libbyolo.so')


def build(force=False, verbose=False):
    cmd = ['make', '-C', CSRC, '-j8'] + (['-B'] if force else [])
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode:
        print(res.stdout)
    if res.returncode:
        raise RuntimeError('libbyolo build failed')
    return LIB


if __name__ == '__main__':
    print(build(verbose=True))
