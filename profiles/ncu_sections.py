"""Attributes the SASS instructions of a raster kernel capture (`ncu --page source --csv --print-source sass,cuda`) to the
sections of the kernel body: size of each section in SASS instructions, how many of them were executed, warp instructions per
camera.  Used for the instruction-cache analysis of DESIGN.md section 9 (the line ranges below are those of the tree the
capture was made from: adjust them to the capture at hand).  Usage: python profiles/ncu_sections.py source.csv"""
import csv,sys,collections
rows=list(csv.reader(open(sys.argv[1])))
hdr=None;cur=None;cur_line=None
addr_lines=collections.defaultdict(list); addr_inst={}
for r in rows:
    if len(r)>=2 and r[0]=="File Path": cur=r[1].split('/')[-1]; continue
    if r and r[0]=="Line No": hdr=r; continue
    if not hdr or len(r)!=len(hdr): continue
    if r[0]!="": cur_line=(cur,int(r[0])); continue
    try: a=int(r[2],16)
    except: continue
    addr_lines[a].append(cur_line); addr_inst[a]=float(r[hdr.index("Instructions Executed")].replace(',','') or 0)
K='raster_kernel.cuh'
sections=[(390,482,'prologue+camera'),(483,551,'dyn list'),(552,572,'segment logic'),(573,612,'flush strips->faces(+quad opt)'),(613,620,'stage1 common'),(621,708,'stage 1S strips'),(709,766,'stage 1F faces'),(767,790,'queue push'),(791,802,'stage2 quad'),(803,816,'stage2 inside'),(817,832,'clipped'),(833,840,'redo'),(841,940,'resolve')]
def sec(a):
    # outermost: a line in the kernel body (>=390)
    for fl,ln in addr_lines[a]:
        if fl==K and ln>=390:
            for lo,hi,name in sections:
                if lo<=ln<=hi: return name
            return 'kernel other'
    return 'unattributed:'+str(addr_lines[a][0])
size=collections.Counter(); execd=collections.Counter(); inst=collections.Counter()
for a in addr_lines:
    s=sec(a); size[s]+=1; inst[s]+=addr_inst[a]
    if addr_inst[a]>0: execd[s]+=1
print('total SASS',len(addr_lines),'executed',sum(execd.values()))
for s,n in size.most_common(): print(f"{s:40s} sass {n:5d} executed {execd[s]:5d}  warp-instr/cam {inst[s]/65536:8.0f}")
