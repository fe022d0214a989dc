"""Registers / stack / spills per kernel from a `-Xptxas -v` log: python scripts/ptxas_summary.py build/hm_wavefront.ptxas.log"""
import re
import sys

txt = open(sys.argv[1]).read()
for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n(?:.*\n)*?ptxas info\s+: Used (\d+) registers[^\n]*", txt):
    blk = m.group(0)
    sp = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", blk)
    name = re.search(r"\d+(k_[a-z_0-9]+)", m.group(1))
    print(f"{(name.group(1) if name else m.group(1)[:40]):28s} regs {m.group(2):>4s}  stack/spill-st/spill-ld {sp.groups() if sp else None}")
