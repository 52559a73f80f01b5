"""SASS evidence of the Blackwell-native path: per kernel of libb200ppo.so, how often the tcgen05 / TMA mnemonics appear.

    python profiles/sass_summary.py > profiles/sass_summary.txt          (no GPU needed: cuobjdump -sass)

UTCHMMA = tcgen05.mma (kind::f16), LDTM = tcgen05.ld, UTMALDG / UTMASTG = cp.async.bulk.tensor load / store,
UBLKCP = cp.async.bulk, SYNCS = mbarrier ops, UTCBAR = tcgen05.commit (B200_PROFILING.md "What proves a Blackwell-native kernel").
"""
import collections
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(os.path.dirname(HERE), "mujoco_reinforcement_learning_b200", "libb200ppo.so")
OPS = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "MUFU.TANH", "HMMA"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
cur, counts, size = None, collections.defaultdict(collections.Counter), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        size[cur] += 1
        for op in OPS:
            if re.search(r"\b" + re.escape(op) + r"\b", line) or (op == "MUFU.TANH" and "MUFU.TANH" in line):
                counts[cur][op] += 1
print(f"{'kernel':64s} {'instr':>6s} " + " ".join(f"{o:>9s}" for o in OPS))
for k in sorted(size, key=lambda k: -counts[k]["UTCHMMA"] * 100000 - size[k]):
    if sum(counts[k].values()) == 0 and not k.startswith("b200ppo::g"):
        continue
    print(f"{k[:64]:64s} {size[k]:6d} " + " ".join(f"{counts[k][o]:9d}" for o in OPS))
