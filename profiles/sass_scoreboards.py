"""Scoreboard assignment of the global loads of a kernel, decoded from the SASS control codes (cuobjdump -sass): how many LDGs name each of the six
scoreboards as their write barrier, and how many branches the kernel has.  usage: python profiles/sass_scoreboards.py file.o kernel-substring"""
import re,sys,subprocess
from collections import Counter
obj,pat=sys.argv[1],sys.argv[2]
txt=subprocess.run(['cuobjdump','-sass',obj],capture_output=True,text=True).stdout
cur=None; funcs={}
for l in txt.splitlines():
    m=re.search(r'Function : (\S+)',l)
    if m: cur=m.group(1); funcs[cur]=[]; continue
    if cur: funcs[cur].append(l)
for name,lines in funcs.items():
    if pat not in name: continue
    ins=[];i=0
    while i<len(lines):
        m=re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/',lines[i])
        if m and i+1<len(lines):
            m2=re.match(r'\s+/\* (0x[0-9a-f]+) \*/',lines[i+1])
            if m2:
                hi=int(m2.group(1),16); ctrl=(hi>>41)&0x7fffff
                ins.append((m.group(2),ctrl&0xf,(ctrl>>5)&7,(ctrl>>8)&7,(ctrl>>11)&0x3f)); i+=2; continue
        i+=1
    c=Counter(x[2] for x in ins if x[0].lstrip('@!P0123456789 ').startswith('LDG'))
    br=sum(1 for x in ins if 'BRA' in x[0])
    print(name[:110], 'instr',len(ins),'LDG by write-barrier',dict(c),'branches',br)
