import sys, time, ctypes as C; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, bsw_b200 as B
E = B.emu_lib()
E.bsw_emu_host_phases.argtypes=[C.POINTER(B.Params)]+[C.c_void_p]*6+[C.c_size_t, C.c_int, C.c_size_t, C.c_void_p]
n=1000000
t=B.synth_tasks("cfg2_150bp", n)
p=B.make_params()
for th,ch in ((1,16384),(2,16384),(4,16384),(8,16384),(12,16384),(16,16384)):
  for rep in range(3):
    ms=np.zeros(8)
    rc=E.bsw_emu_host_phases(C.byref(p), t['qbuf'].ctypes.data,t['qoff'].ctypes.data,t['tbuf'].ctypes.data,t['toff'].ctypes.data,t['h0'].ctypes.data,t['w'].ctypes.data,n,th,ch,ms.ctypes.data)
    print(th, ch, "fill %.1f validate+pack %.1f plan %.1f - %.1f (cpu-ms summed) wall %.1f"%tuple(ms[:5]))
