"""Side-by-side per-layer times of several tools/profile_layers.py outputs (A/B of kernel variants on one GPU box):
    python tools/compare_layers.py gpurun_out/sweep_x_0.txt gpurun_out/sweep_x_1.txt ...
Rows are the fused conv layers keyed by (channels, kernel size, tensor passes); `floor` = max(bytes / 6550 GB/s,
flops / 1394.7 TFLOP/s) in ms (MEASURED_PEAKS.json)."""
import sys,re,collections
files=sys.argv[1:]
tabs=[]
keys=[]
for f in files:
    d=collections.OrderedDict()
    for line in open(f):
        m=re.match(r'(\S+)\s+(\d+)\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)\s+GF=([\d.]+) MB=([\d.]+)',line)
        if m and m.group(1) in ('conv_fused','conv_pipe','conv_row'):
            k=(float(m.group(8)),float(m.group(9)))
            d[k]=(int(m.group(2)),float(m.group(4)))
            if k not in keys: keys.append(k)
    tabs.append(d)
def cls(gf,mb):
    # derive C and k
    for C,T in ((32,120000),(64,60000),(128,20000),(256,4000)):
        per=2*64*T*C*C/1e9
        k=gf/per
        if abs(k-round(k))<0.02 and round(k) in (3,7,11):
            tb=64*T*C*2/1e6          # passes in units of one fp16 tensor (x, y, residual, old values)
            n=mb/tb
            return C,int(round(k)),round(n)
    return None
rows=[]
for k in keys:
    c=cls(*k)
    rows.append((c if c else (999,0,0),k))
rows.sort()
print('%-22s'%'layer (C,k,fp16 passes) n', ' '.join('%8s'%('cfg%d'%i) for i in range(len(files))), '  floor')
tot=[0]*len(files)
for c,k in rows:
    n=[t.get(k,(0,0))[0] for t in tabs][0]
    floor=max(k[1]/6550.1/1e3*1e3/1e3, k[0]/1394.7/1e3) # ms: MB/(GB/s) -> ms ; GF/(TF/s) -> ms
    floor=max(k[1]/6550.1, k[0]/1394.7)
    print('%-22s'%(str(c)+' x%d'%n), ' '.join('%8.4f'%t.get(k,(0,0))[1] for t in tabs), '  %.3f'%floor)
    for i,t in enumerate(tabs): tot[i]+=t.get(k,(0,0))[0]*t.get(k,(0,0))[1]
print('%-22s'%'conv_fused total ms', ' '.join('%8.2f'%x for x in tot), '  %.2f'%sum(max(k[1]/6550.1,k[0]/1394.7)*tabs[0].get(k,(0,0))[0] for c,k in rows))
