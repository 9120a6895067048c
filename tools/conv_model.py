"""Per-layer traffic / time model of the conv stack (B=16, 1088x1920) vs the ncu launch list.  Usage: python tools/conv_model.py launches.csv"""
import csv, sys, math
sys.path.insert(0, '.')
B=16; H=1088; W=1920
def plan():
    ops=[]
    def conv(name,cin,cout,k,s,h,w,res=False,up=False,f32=False): ops.append(dict(name=name,cin=cin,cout=cout,k=k,s=s,h=h,w=w,res=res,up=up,f32=f32))
    def c2f(pre,cin,c2,n,h,w,sc,up=False):
        c=c2//2
        conv(pre+'.cv1',cin,2*c,1,1,h,w)
        for i in range(n):
            conv(f'{pre}.m{i}.cv1',c,c,3,1,h,w); conv(f'{pre}.m{i}.cv2',c,c,3,1,h,w,res=sc)
        conv(pre+'.cv2',(2+n)*c,c2,1,1,h,w,up=up)
    h1,w1=H//2,W//2; h2,w2=H//4,W//4; h3,w3=H//8,W//8; h4,w4=H//16,W//16; h5,w5=H//32,W//32
    ops.append(dict(name='0',cin=16,cout=32,k=2,s=1,h=h1,w=w1,res=False,up=False,f32=False,kbe=16))
    conv('1',32,64,3,2,h2,w2); c2f('2',64,64,1,h2,w2,True)
    conv('3',64,128,3,2,h3,w3); c2f('4',128,128,2,h3,w3,True)
    conv('5',128,256,3,2,h4,w4); c2f('6',256,256,2,h4,w4,True)
    conv('7',256,512,3,2,h5,w5); c2f('8',512,512,1,h5,w5,True)
    conv('9.cv1',512,256,1,1,h5,w5); conv('9.cv2',1024,512,1,1,h5,w5,up=True)
    c2f('12',768,256,1,h4,w4,False,up=True); c2f('15',384,128,1,h3,w3,False)
    conv('16',128,128,3,2,h4,w4); c2f('18',384,256,1,h4,w4,False)
    conv('19',256,256,3,2,h5,w5); c2f('21',768,512,1,h5,w5,False)
    for i,(c,h,w) in enumerate(((128,h3,w3),(256,h4,w4),(512,h5,w5))):
        conv(f'h{i}.a',c,192,3,1,h,w); conv(f'h{i}.b2',64,64,3,1,h,w); conv(f'h{i}.b3',128,128,3,1,h,w)
        conv(f'h{i}.r2',64,64,1,1,h,w,f32=True); conv(f'h{i}.r3',128,4,1,1,h,w,f32=True)
    return ops
def model(o):
    kbe=o.get('kbe',64)
    tiles=B*math.ceil(o['h']*o['w']/128)  # approx
    cout16=math.ceil(o['cout']/16)*16; BN=min(cout16,256); nt=math.ceil(cout16/BN)
    kc=math.ceil(o['cin']/kbe); nkb=o['k']*o['k']*kc
    a_real=128*min(o['cin'],kbe)*2 if kc==1 else 128*kbe*2
    a_l2=tiles*nt*nkb*a_real
    b_l2=tiles*nt*nkb*BN*kbe*2
    outb=B*o['h']*o['w']*o['cout']*(4 if o['f32'] else 2)*(5 if o['up'] else 1)
    resb=B*o['h']*o['w']*o['cout']*2 if o['res'] else 0
    inb=B*(o['h']*o['s'])*(o['w']*o['s'])*o['cin']*2
    flops=2*B*o['h']*o['w']*o['cout']*o['cin']*o['k']**2
    mma_cyc=tiles*nt*nkb*(kbe//16)*max(128*BN/256,1)   # cycles summed over tiles
    return dict(tiles=tiles*nt,nkb=nkb,a_l2=a_l2,b_l2=b_l2,hbm=inb+outb+resb,flops=flops,mma_us=mma_cyc/148/1.9e3,out=outb)
if __name__=='__main__':
    ops=plan()
    meas=None
    if len(sys.argv)>1:
        rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10 and r[0].isdigit()]
        meas=[float(r[-1])/1e3 for r in rows if 'conv_tc' in r[4]][:len(ops)]
    tot=dict(t=0,l2=0,hbm=0,mma=0,rf=0)
    print(f"{'op':10s} {'tiles':>6s} {'nkb':>3s} {'A_L2MB':>8s} {'B_L2MB':>8s} {'HBM_MB':>8s} {'t_l2':>6s} {'t_hbm':>6s} {'t_mma':>6s} {'meas':>7s}")
    for i,o in enumerate(ops):
        m=model(o)
        t_l2=(m['a_l2']+m['b_l2']+m['out'])/6.5e6; t_hbm=m['hbm']/6.4e6
        ms=meas[i] if meas else float('nan')
        tot['t']+=ms; tot['l2']+=t_l2; tot['hbm']+=t_hbm; tot['mma']+=m['mma_us']; tot['rf']+=max(t_hbm,m['flops']/1413e6)
        print(f"{o['name']:10s} {m['tiles']:6d} {m['nkb']:3d} {m['a_l2']/1e6:8.0f} {m['b_l2']/1e6:8.0f} {m['hbm']/1e6:8.0f} {t_l2:6.0f} {t_hbm:6.0f} {m['mma_us']:6.0f} {ms:7.0f}")
    print(tot)
