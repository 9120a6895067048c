"""Source-line view of an `ncu --set full --import-source on` capture without the GUI: the SASS rows of `--page source --csv` (instructions
executed, stall samples per instruction) are mapped to CUDA source lines through `nvdisasm -g` line info of the object the kernel was
built from (the library is compiled with -lineinfo), and the lines are ranked by stall samples.

    ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:fast_kernel -o gpurun_out/k python tools/profile_step.py 1
    python tools/ncu_source_lines.py gpurun_out/k.ncu-rep fast_kernel geo-trax_b200/build/orb.o geo-trax_b200/csrc/orb.cu [top]

This is how round 2 found the byte-wise patch staging of orb_describe_kernel (41 % of its stall samples) and the flat profile of fast_kernel."""
import csv,re,collections,subprocess,sys,os
rep, kern, objfile, srcfile = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 14
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--kernel-name',f'regex:{kern}'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.split('\n')))
his=[i for i,r in enumerate(rows) if r and r[0]=='Address']
hi=his[0]; end = his[1]-1 if len(his)>1 else len(rows)
h=rows[hi]; ci={n:i for i,n in enumerate(h)}
def f(x):
    try: return float(x)
    except: return 0.0
sass=[(r[ci['Source']].strip(), f(r[ci['Instructions Executed']]), f(r[ci['# Samples']])) for r in rows[hi+1:end] if len(r)>ci['Instructions Executed']]
d='/tmp/t/dis_'+os.path.basename(objfile)+'.txt'
if not os.path.exists(d):
    os.makedirs('/tmp/t/x',exist_ok=True)
    subprocess.run(f'cd /tmp/t/x && rm -f *.cubin && cuobjdump -xelf all {objfile} >/dev/null 2>&1 && nvdisasm -g -c $(ls *.cubin|head -1) > {d}',shell=True)
lines=open(d).read().split('\n')
start=next(i for i,l in enumerate(lines) if l.startswith('.text.') and re.search(kern,l))
cur=None; seq=[]
for l in lines[start+1:]:
    if l.startswith('//---------------------'): break
    m=re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur=(m.group(1).split('/')[-1], int(m.group(2))); continue
    m=re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: seq.append((cur, m.group(2).strip()))
per=collections.Counter(); samp=collections.Counter()
if len(sass)!=2*len(seq) and len(sass)!=len(seq): print("WARNING length mismatch", len(sass), len(seq))
for i in range(min(len(sass),len(seq))):
    per[seq[i][0]]+=sass[i][1]; samp[seq[i][0]]+=sass[i][2]
tot=sum(per.values()) or 1; ts=sum(samp.values()) or 1
src=open(srcfile).read().split('\n')
print(f"== {kern}: {tot:.4g} warp instr, {ts:.0f} samples")
for loc,c in samp.most_common(top):
    fn,ln=loc if loc else ('?',0)
    text=src[ln-1].strip()[:105] if fn==os.path.basename(srcfile) and ln-1<len(src) else fn
    print(f"{fn}:{ln:4d} {100*per[loc]/tot:5.1f}% instr {100*c/ts:5.1f}% samp | {text}")
