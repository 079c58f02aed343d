import sys, importlib, struct
sys.path.insert(0,'/root/repo')
import numpy as np
import __graft_entry__ as g; g.build()
mic=importlib.import_module('medical-image-codec_b200'); synth=importlib.import_module('medical-image-codec_b200.synth')
from oracle.oracle import Oracle
o=Oracle()
W,H=301,203
rgb = synth.wsi_region(5, 1000, 900, W, H, 2500, 2000)
blob = o.rgb_compress(rgb.ravel(), W, H, True)
l=struct.unpack_from('<3I',blob,0); off=12
y,co,cg=o.ycocg_forward(rgb.ravel())
for k,(ln,ref) in enumerate(zip(l,(y,co,cg))):
    pb=blob[off:off+ln]; off+=ln
    print('plane',k,'mode',pb[0],'len',ln, 'max',ref.max())
    if pb[0]==2:
        got=mic.DecompressSingleFrame(pb[1:],W,H)
        bad=np.nonzero(got!=ref)[0]
        print('  mismatches',bad.size, 'first',bad[:5], [(int(i)//W,int(i)%W) for i in bad[:5]])
        if bad.size:
            i=int(bad[0]); print('  got',got[i-2:i+6],'ref',ref[i-2:i+6])
        sym=o.delta_rle_compress(ref,W,H,max(int(ref.max()),255)); print('  nsym',sym.size,'ndelim',int((sym==sym[0]).sum()), 'tableinfo', o.fse_table_info(sym)[:2])
got=mic.DecompressRGB(blob,W,H)
bad=np.nonzero(got!=rgb.ravel())[0]; print('rgb mismatches',bad.size, bad[:10])
