"""Per-source-line executed warp instructions and stall samples of one kernel of an `ncu --set full --import-source on` report:
the SASS page of the report is joined with nvdisasm line info of the object the kernel was built from (-lineinfo).
usage: [SRCDIR=dir/with/the/sources/] python tools/ncu_lines.py report.ncu-rep object.o kernel_name_substring"""
import csv,sys,subprocess,io,collections,re,os,glob,tempfile
rep,obj,kname=sys.argv[1:4]
src=subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(src)))
hdr=None
for i,r in enumerate(rows):
    if "Source" in r and "# Samples" in r:
        hdr=r; start=i+1; break
ci=hdr.index('# Samples'); si=hdr.index('Source'); ie=hdr.index('Instructions Executed'); ia=hdr.index('Address')
data=[r for r in rows[start:] if len(r)>ci and r[ia].startswith('0x')]
base=int(data[0][ia],16)
prof={}
for r in data:
    prof[int(r[ia],16)-base]=(int(r[ie] or 0),int(r[ci] or 0),r[si].strip())
d=tempfile.mkdtemp()
subprocess.run(["cuobjdump","-xelf","all",os.path.abspath(obj)],cwd=d,capture_output=True)
cub=glob.glob(d+"/*.cubin")[0]
dis=subprocess.run(["nvdisasm","-g","-c",cub],capture_output=True,text=True).stdout.split('\n')
# find the function
infn=False; line=None; off2line={}
mangled=None
for ln in dis:
    m=re.match(r'\s*\.section\s+\.text\.(\S+),',ln)
    if m:
        infn = kname in m.group(1)
        continue
    if not infn: continue
    m=re.search(r'//## File "([^"]+)", line (\d+)',ln)
    if m: line=(os.path.basename(m.group(1)),int(m.group(2))); continue
    m=re.match(r'\s*/\*([0-9a-f]{4,})\*/',ln)
    if m: off2line[int(m.group(1),16)]=line
agg=collections.defaultdict(lambda:[0,0])
toti=tots=0
for off,(ne,ns,s) in prof.items():
    l=off2line.get(off,('?',0))
    agg[l][0]+=ne; agg[l][1]+=ns; toti+=ne; tots+=ns
print(f"total warp-inst {toti}  samples {tots}  mapped offsets {len(off2line)} / profiled {len(prof)}")
srcs={}
for (f,l),(ne,ns) in sorted(agg.items(), key=lambda kv:(kv[0][0],kv[0][1])):
    if ne*200<toti and ns*200<tots: continue
    if f not in srcs:
        p=[q for q in glob.glob(os.environ.get('SRCDIR','/root/repo/structured-alignment-vqa_b200/csrc/')+f)]
        srcs[f]=open(p[0]).read().split('\n') if p else []
    text=srcs[f][l-1].strip()[:100] if srcs[f] and l-1<len(srcs[f]) else ''
    print(f"{f}:{l:4d} inst {100*ne/toti:5.1f}% smp {100*ns/max(tots,1):5.1f}% | {text}")
