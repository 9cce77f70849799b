"""Lane-exact Python model of the batch rules of lz4_decode_spec.cu (K1 for word-regular columns).

Test infrastructure: an executable statement of WHAT the kernel does with a raw LZ4 block -- which sequences a batch takes
(full batch of 32 one-word sequences, full batch with one two-word or one (1, 7) sequence, run + closing sequence), how the
stream position and the run's shape (L0) move, which source a word is fetched from (ring / global memory) and how in-batch
chains collapse by pointer jumping over the word index -- checked on the CPU against the oracle's codec
(tests/test_decoder_models.py).  The CUDA kernel itself is checked against the oracle on the GPU (tests/test_gpu_parity.py);
the rules restated here follow read_block / LZ4_decompress_safe of the reference (src/io/BlockStreams.jl:101-119)."""

RING = 512
NEAR = RING - 64
M64 = (1 << 64) - 1


def decode_block_spec(src, origin):
    n_src=len(src); src=bytes(src)+bytes(4096)
    out=bytearray(origin)
    ring=[None]*RING
    def w64(i): return int.from_bytes(out[8*i:8*i+8],'little')
    ip=op=0; L0=0; pendL=0xff; tok0=4; sh0=8; kp0=(1<<64)-1; ring_from=0
    lim_w=(origin-12)>>3 if origin>=12 else 0
    ip_lim=n_src-240 if n_src>=240 else 0
    fast=n_src>=240
    M64=(1<<64)-1
    tp=[3*l for l in range(32)]
    def ld(t): return int.from_bytes(src[t:t+8],'little')
    x=[ld(t) for t in tp]
    stats={'fb':0,'nfb':0,'one':0}
    def source(sw,opw):
        if sw>=ring_from and opw-sw<=NEAR:
            v=ring[sw&(RING-1)]
            assert v is not None and v==w64(sw), (sw,opw,v,w64(sw))
            return v
        return w64(sw)
    def put(w,v):
        out[8*w:8*w+8]=v.to_bytes(8,'little'); ring[w&(RING-1)]=v
    done=False
    while not done:
        batch=False
        if fast and (op&7)==0 and ip<=ip_lim:
            opw=op>>3
            tok=[xx&0xff for xx in x]; off=[(xx>>sh0)&0xffff for xx in x]
            offr=[((o>>3)|(o<<29))&0xffffffff for o in off]
            okp=[tok[l]==tok0 and ((offr[l]-1)&0xffffffff)<opw+l and opw+l<lim_w for l in range(32)]
            bad=[l for l in range(32) if not okp[l]]
            fb1=False
            if len(bad)==1:
                z=bad[0]
                fb1=(tok[z]==tok0+8 and ((offr[z]-1)&0xffffffff)<opw+z and offr[z]>=z+2 and opw+33<=lim_w)
            if fb1:
                stride=3+L0; pendL=0xff
                nip=ip+32*stride; tp=[t+32*stride for t in tp]; nx=[ld(t) for t in tp]
                pos=[l+(1 if l>z else 0) for l in range(32)]
                s=[opw+pos[l]-offr[l] for l in range(32)]
                inb=[s[l]>=opw for l in range(32)]
                assert not inb[z]
                while any(inb):
                    snap=list(s)
                    for l in range(32):
                        q=s[l]-opw
                        if inb[l]:
                            own=q-(1 if q>z else 0)
                            s[l]=snap[own]+(1 if q==z+1 else 0); inb[l]=s[l]>=opw
                vals=[(source(s[l],opw)&kp0)|((x[l]>>8)&~kp0&M64) for l in range(32)]
                v2=source(s[z]+1,opw)
                for l in range(32): put(opw+pos[l],vals[l])
                put(opw+pos[z]+1,v2)
                ip=nip; op+=8*33; x=nx; batch=True; stats['fb1']=stats.get('fb1',0)+1
                continue
            fb1b=False
            if not all(okp) and L0==0 and tok0==4:
                z=okp.index(False)
                tok1=[(xx>>8)&0xff for xx in x]; off1=[(xx>>16)&0xffff for xx in x]
                offr1=[((o>>3)|(o<<29))&0xffffffff for o in off1]
                cond=[]
                for l in range(32):
                    if l<z: cond.append(True)
                    elif l==z: cond.append(tok[l]==0x13 and ((offr1[l]-1)&0xffffffff)<opw+l and opw+l<lim_w)
                    else: cond.append(tok1[l]==4 and ((offr1[l]-1)&0xffffffff)<opw+l and opw+l<lim_w)
                refz=any(l>z and opw+l-offr1[l]==opw+z for l in range(32))
                fb1b=all(cond) and not refz
            if fb1b:
                pendL=0xff
                nip=ip+97; tp=[t+97 for t in tp]; nx=[ld(t) for t in tp]
                oe=[offr1[l] if l>=z else offr[l] for l in range(32)]
                s=[opw+l-oe[l] for l in range(32)]
                inb=[s[l]>=opw for l in range(32)]
                while any(inb):
                    snap=list(s)
                    for l in range(32):
                        if inb[l]: s[l]=snap[(s[l]-opw)&31]; inb[l]=s[l]>=opw
                vals=[source(s[l],opw) for l in range(32)]
                vals[z]=(vals[z]&~0xff&M64)|((x[z]>>8)&0xff)
                for l in range(32): put(opw+l,vals[l])
                ip=nip; op+=256; x=nx; batch=True; stats['fb1b']=stats.get('fb1b',0)+1
                continue
            if all(okp):
                n=32;W_s=0;hdr_s=0;srcw=offr;kp=[kp0]*32;sp=[False]*32;stride=3+L0;pendL=0xff;FB=True;go=True
            else:
                FB=False
                n=okp.index(False)
                sp=[False]*32;W_s=0;hdr_s=0
                l=n;L=tok[l]>>4;LM=L+(tok[l]&15)+4
                off_s=(x[l]>>((8+8*L)&63))&0xffff;offw_s=off_s>>3;W=LM>>3;myw=opw+l
                if L<=5 and (tok[l]&15)!=15 and (LM&7)==0 and (off_s&7)==0 and off_s!=0 and offw_s<=myw and offw_s>=l+W and myw+W<=lim_w:
                    sp[l]=True;hdr_s=3+L;W_s=W
                go=n+W_s>0
                if go:
                    stride=3+L0
                    if n==0 and W_s==1:
                        Lh=hdr_s-3
                        if Lh==pendL and Lh<=4:
                            L0=Lh;tok0=(L0<<4)|(4-L0);sh0=8+8*L0;kp0=(M64<<(8*L0))&M64
                        pendL=Lh
                    else: pendL=0xff
                    srcw=[offw_s if sp[l] else offr[l] for l in range(32)]
                    kp=[((M64<<(8*L))&M64) if sp[l] else kp0 for l in range(32)]
            if go:
                adv=32*stride if FB else stride*n+hdr_s
                nip=ip+adv
                if (not FB) and stride!=3+L0: tp=[nip+(3+L0)*l for l in range(32)]
                else: tp=[t+adv for t in tp]
                assert tp==[nip+(3+L0)*l for l in range(32)]
                nx=[ld(t) for t in tp]
                s=[(opw+l-srcw[l]) for l in range(32)]
                mine=[FB or l<n for l in range(32)]
                inb=[mine[l] and s[l]>=opw for l in range(32)]
                while any(inb):
                    t=[s[(s[l]-opw)&31] for l in range(32)]
                    for l in range(32):
                        if inb[l]: s[l]=t[l]; inb[l]=s[l]>=opw
                vals={}
                for l in range(32):
                    if mine[l] or sp[l]:
                        v=source(s[l],opw); vals[l]=(v&kp[l])|((x[l]>>8)&~kp[l]&M64)
                for l,v in vals.items(): put(opw+l,v)
                for l in range(32):
                    if W_s==2 and sp[l]: put(opw+l+1,source(s[l]+1,opw))
                ip=nip; op+=8*(32 if FB else n+W_s); x=nx; batch=True
                stats['fb' if FB else 'nfb']+=1
        if not batch:
            stats['one']+=1
            tok=src[ip];ip+=1;L=tok>>4
            if L==15:
                while True:
                    e=src[ip];ip+=1;L+=e
                    if e!=255:break
            out[op:op+L]=src[ip:ip+L];ip+=L;op+=L
            if ip==n_src: done=True
            else:
                off=src[ip]|(src[ip+1]<<8);ip+=2;M=tok&15
                if M==15:
                    while True:
                        e=src[ip];ip+=1;M+=e
                        if e!=255:break
                M+=4
                for i in range(M): out[op+i]=out[op-off+i]
                op+=M
            ring_from=(op+7)>>3
            tp=[ip+(3+L0)*l for l in range(32)]
            x=[ld(t) for t in tp]
    return bytes(out),stats


