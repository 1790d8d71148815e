import sys; sys.path.insert(0,'tests')
import bpp
e = bpp.engine()
names = {0:'IMAD.lo',1:'IMAD.HI',2:'IMAD.WIDE',3:'IADD+LOP',10:'IMAD+IADD',11:'IMAD.WIDE+IADD',4:'fe_mul',5:'fe_sq',9:'fe_mul_portable',6:'madd',7:'dbl',8:'sc_montmul'}
for w,nm in names.items():
    it = 4000 if w in (0,1,2,3,10,11) else 400
    ops,sec = e.microbench(w,it)
    print("%-16s %.4e ops/s  (%.3f ms)"%(nm,ops,sec*1e3))
