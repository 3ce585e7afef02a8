import re,collections,subprocess,sys
F='k_sweep_tma_psiIfLi1ELi6ELi2ELi2ELb1ELi640'      # substring of the mangled name of the production instantiation
raw=subprocess.run(['cuobjdump','-sass','semiclassicalmc.jl_b200/libscmc_b200.so'],capture_output=True,text=True).stdout
raw=[b for b in raw.split('Function : ') if F in b.split('\n')[0]][0]
lines=[re.sub(r'\s+/\* 0x[0-9a-f]+ \*/','',l) for l in raw.splitlines() if re.match(r'^\s+/\*[0-9a-f]{4}\*/',l)]
def addr(l): return int(re.search(r'/\*([0-9a-f]{4})\*/',l).group(1),16)
def op(l):
    m=re.search(r'\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)',l); return m.group(1) if m else '?'
best=None
for l in lines:
    m=re.search(r'@!?U?P\d\s+BRA 0x([0-9a-f]+)',l)      # conditional backward branch (the out-of-line retry stubs jump back unconditionally)
    if m:
        t=int(m.group(1),16); a=addr(l)
        if t<a and (best is None or a-t>best[1]-best[0]): best=(t,a)
lo,hi=best
loop=[l for l in lines if lo<=addr(l)<=hi]
out=[]
out.append("k_sweep_tma_psi<float, GEN=1, Z=6, STREAM=2, HITS=2, TC=true, CAPC=640> -- SASS of the production kernel of the headline workload (sm_100a, `cuobjdump -sass` of the")
out.append("in-tree build, round 2, third session; regenerate with the script at the end of profiles/README.md; second session: r02_sass_k_sweep_tma_psi_tc_session2.txt)")
out.append(f"replica loop: {lo:#06x} .. {hi:#06x} (back edge), {len(loop)} static instructions; ncu counts 659 executed warp-instructions per warp and replica for the whole")
out.append("kernel (the MMA issue block runs on one warp of four, the bulk-copy issue block on another; prologue / epilogue amortised over 64 replicas per CTA).")
out.append("")
out.append("Blackwell-specific instructions in the kernel (PTX -> SASS: tcgen05.alloc/dealloc -> UTCATOMSWS, tcgen05.st/ld -> STTM/LDTM, tcgen05.mma.kind::tf32 -> UTCHMMA,")
out.append("tcgen05.commit -> UTCBAR, cp.async.bulk -> UBLKCP, mbarrier.* -> SYNCS.*, elect.sync -> ELECT):")
for l in lines:
    if re.search(r'UTCHMMA|UTCBAR|STTM|LDTM|UTCATOMSWS|UBLKCP|SYNCS\.|FENCE\.VIEW|ELECT',l): out.append("  "+l.strip())
c=collections.Counter(op(l) for l in loop)
groups=collections.OrderedDict([('FFMA2  packed density sums of six neighbours (96), three energies as psi^dagger (H psi) (3 x 16), two proposals (2 x 8)',0),('FFMA',0),('FMUL / FADD / FMUL2',0),('IMAD.WIDE  Philox4x32-7: 3 calls per site visit',0),('MUFU   2 x (4 LG2 + 4 SQRT + 4 SIN + 4 COS) Box-Muller, 2 RSQ, 2 EX2',0),('LDS    14 x 128-bit staged states (immediate plane / stage offsets) + per-replica scalars',0),('tensor core / tensor memory / bulk copy / mbarrier (UTCHMMA x 6, UTCBAR, STTM x 2, LDTM, UBLKCP, SYNCS, ELECT, FENCE)',0),('integer ALU (LOP3 / SHF / LEA / IADD3 / PRMT / IMAD / I2F / VOTE / POPC / ISETP ...)',0),('other (MOV / BRA / BSSY / uniform datapath / STG / RED / NOP)',0)])
keys=list(groups)
for k,v in c.items():
    if k.startswith('FFMA2'): groups[keys[0]]+=v
    elif k.startswith('FFMA'): groups[keys[1]]+=v
    elif k.startswith('FMUL') or k.startswith('FADD'): groups[keys[2]]+=v
    elif k.startswith('IMAD.WIDE'): groups[keys[3]]+=v
    elif k.startswith('MUFU'): groups[keys[4]]+=v
    elif k.startswith('LDS'): groups[keys[5]]+=v
    elif re.match(r'UTC|STTM|LDTM|UBLKCP|SYNCS|FENCE|ELECT',k): groups[keys[6]]+=v
    elif re.match(r'LOP3|SHF|LEA|IADD3|PRMT|IMAD|VIADD|I2F|POPC|VOTE|ISETP|SEL|PLOP3',k): groups[keys[7]]+=v
    else: groups[keys[8]]+=v
out.append("")
out.append("instruction histogram of the replica loop (static, one pass; second session: 852 instructions, 136 FFMA + 96 FFMA2 + 146 FMUL / FADD):")
for k,v in groups.items(): out.append(f"  {v:5d}  {k}")
out.append("")
out.append("full opcode counts:")
for k,v in c.most_common(): out.append(f"  {v:5d}  {k}")
open('profiles/r02_sass_k_sweep_tma_psi_tc.txt','w').write("\n".join(out)+"\n")
