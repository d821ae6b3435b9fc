#!/usr/bin/env python3
"""profiles/rNN_sass_hot_kernels.txt: opcode histograms and Blackwell-specific instructions of the hot
kernels, from `cuobjdump -sass sfft_b200/build/*.o` (no GPU needed).  usage: tools/sass_evidence.py > file"""
import collections
import re
import subprocess

OBJS = {
    "v12_kernels.cu.o": ["gather_kernelILb1ELi2E", "gather_kernelILb1ELi8E", "v2_fused_kernelILi20E", "v2_regroup_kernel",
                         "select_kernel", "select_cluster_kernel", "vote_kernelILb1ELb0E", "vote_kernelILb1ELb1E",
                         "estimate_pair_kernelILi20E"],
    "v3.cu.o": ["v3_peel_kernel"],
    "shard.cu.o": ["shard_push_kernel", "shard_wait_ready_kernel", "shard_done_kernel"],
    "fft.cu.o": ["fft_pass_kernelILi1E"],
}
MARK = re.compile(r"UBLKCP|SYNCS|UCGABAR|CCTL|MEMBAR|ERRBAR|ATOMS|REDS|ATOMG|REDG|LDG.*(NA|LTC|CONSTANT)|ST\.E.*SYS|STG.*SYS|LDG.*SYS|"
                  r"LD\.E.*SYS|LDS\.128|STS\.128|MAPA|UTMA|NANOSLEEP")

print("# SASS evidence for the hot kernels (cuobjdump -sass of sfft_b200/build/*.o, sm_100a; tools/sass_evidence.py)")
print("# per kernel: opcode histogram (top 14) and the instructions that show the Blackwell-specific paths:")
print("#   UBLKCP = cp.async.bulk (TMA engine), SYNCS = mbarrier, UCGABAR_* = barrier.cluster, LDG...LTC64B = .L2::64B fills,")
print("#   ATOMS/REDS/MAPA = shared-memory atomics incl. distributed shared memory, *.SYS + MEMBAR.*.SYS = the peer-memory")
print("#   protocol of the sharded transform (stores into CUDA-IPC-mapped peer buffers, flags with release/acquire at system scope)")
print("#   v2_fused_kernel: FMNMX/FMNMX3 = the medians on 32-bit keys (hot path); its FSEL/DSETP are the exact 64-bit networks of the")
print("#   tie / out-of-band fallback (cold)")
print()
for obj, kerns in OBJS.items():
    sass = subprocess.run(["cuobjdump", "-sass", "sfft_b200/build/" + obj], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)
    for k in kerns:
        for f in funcs:
            name = f.split("\n", 1)[0]
            if k in name:
                ops = re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f, re.M)
                hist = collections.Counter(o.split(".")[0] for o in ops)
                special = collections.Counter(o for o in ops if MARK.search(o))
                print(f"== {name[:120]}")
                print(f"   instructions {len(ops)}; " + ", ".join(f"{a} {b}" for a, b in hist.most_common(14)))
                if special:
                    print("   marks: " + ", ".join(f"{a} x{b}" for a, b in sorted(special.items())))
                print()
                break
