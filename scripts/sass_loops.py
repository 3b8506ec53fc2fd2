#!/usr/bin/env python
"""Instruction mix of the loops of one kernel, from `cuobjdump -sass` (no GPU needed):

    python scripts/sass_loops.py moldyn_b200/lib/libmoldyn_b200.so k_forceILb0ELi2ELb1ELb0 100

prints, for every backward branch spanning more than <min_len> instructions, the loop's length and opcode histogram, plus the
kernel's local-memory traffic (STL/LDL = spills).  This is how the dense loop's if-converted potential/virial arithmetic and
its register-rotation moves were found."""
import re,sys,subprocess
from collections import Counter
lib=sys.argv[1]; pat=sys.argv[2]
minlen=int(sys.argv[3]) if len(sys.argv)>3 else 20
txt=subprocess.run(["cuobjdump","-sass",lib],capture_output=True,text=True).stdout
funcs=re.split(r'\n\s+Function : ',txt)
for f in funcs[1:]:
    name=f.split('\n')[0]
    if pat not in name: continue
    print("==",name[:90])
    ins=[]
    for ln in f.split('\n'):
        m=re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);',ln)
        if m: ins.append((int(m.group(1),16),m.group(2).strip()))
    idx={a:i for i,(a,_) in enumerate(ins)}
    print("total instr",len(ins),"STL",sum('STL' in x for _,x in ins),"LDL",sum('LDL' in x for _,x in ins))
    for i,(a,t) in enumerate(ins):
        m=re.search(r'BRA\s+(?:\w+,\s*)?0x([0-9a-f]+)',t)
        if m:
            tgt=int(m.group(1),16)
            if tgt<a and (a-tgt)//16>minlen:
                body=ins[idx[tgt]:i+1]
                c=Counter()
                for _,x in body:
                    x=re.sub(r'^@!?U?P\d+\s+','',x)
                    c[x.split()[0].split('.')[0]]+=1
                print(hex(tgt),hex(a),'n',len(body),dict(c.most_common(16)))
