"""SASS opcode census of the step kernels (cuobjdump -sass of the built library): which instructions prove the
tcgen05 / TMEM / bulk-copy path.  Usage: python profiles/sass_census.py > profiles/round2_sass_census.txt"""
import collections
import re
import subprocess
import sys

LIB = "trajnetplusplusbaselines_b200/libtrajnet_b200.so"
KEEP = ("pool_prepare", "sparse_layer1_pair", "sparse_layer1_tc", "dense_layer_tc", "lstm_gates_tc", "social_dgrid_mma",
        "hidden_mlp_pool", "sf_simulate", "orca_simulate", "scenes_", "traj_", "pool_lstm_cell")
OPS = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UBLKCP", "SYNCS", "UCGABAR_ARV", "UCGABAR_WAIT",
       "HMMA", "LDGSTS", "MATCH", "FFMA", "DFMA", "DMUL", "DADD", "F2F", "MUFU", "MEMBAR", "CCTL", "ACQBULK", "UTCCP")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
print("SASS opcode census of the step kernels in libtrajnet_b200.so (cuobjdump -sass, sm_100a), round 2 final")
print("tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA tensor loads -> UTMALDG, cp.async.bulk -> UBLKCP, tcgen05.commit -> UTCBAR,")
print("mbarrier -> SYNCS, cluster barrier -> UCGABAR, cp.async -> LDGSTS; HMMA is the warp-level mma.sync path.\n")
name, counts, total = None, None, 0


def flush():
    if name and any(k in name for k in KEEP):
        print("%s  (%d instructions)" % (name, total))
        for op, n in sorted(counts.items(), key=lambda kv: -kv[1]):
            print("    %-28s %d" % (op, n))
        print()


for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush()
        name, counts, total = m.group(1), collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        total += 1
        op = m.group(1).split(".")[0]
        if op in OPS:
            counts[op] += 1
flush()
