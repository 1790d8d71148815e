import sys; sys.path.insert(0,'tests')
import bpp
e = bpp.engine()
names = {0:'IMAD.lo',1:'IMAD.HI',2:'IMAD.WIDE',3:'IADD+LOP',10:'IMAD+IADD',11:'IMAD.WIDE+IADD',4:'fe_mul',5:'fe_sq',9:'fe_mul_portable',6:'madd',7:'dbl',8:'sc_montmul'}
for w,nm in names.items():
    it = 4000 if w in (0,1,2,3,10,11) else 400
    ops,sec = e.microbench(w,it)
    print("%-16s %.4e ops/s  (%.3f ms)"%(nm,ops,sec*1e3))
print("--- latency probes: one warp per SM, dependent chain; ns per op")
lat = {20:'fe_mul',21:'fe_sq',24:'sc_montmul',25:'ge_dbl',27:'ge_madd',30:'quad_dbl',31:'quad_add',32:'quad_add+to_cached'}
for w,nm in lat.items():
    it = 2000
    ops,sec = e.microbench(w,it)
    print("%-16s %.1f ns/op"%(nm, sec/(2*it)*1e9))
