"""Dry run of bench.py's main() over a stand-in torch.cuda and a stand-in ctx (tests only): catches
name / key errors in the host logic of the bench (JSON assembly, roofline table, e2e, cpu_baseline,
ride-along configs) without a GPU.  Run by tests/test_bench_contract.py."""
import sys, types, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); os.chdir(ROOT)
import numpy as np
# ---- fake torch
torch = types.ModuleType('torch')
class _Cuda:
    def set_device(self, d): pass
    def synchronize(self): pass
    def is_available(self): return True
    def device_count(self): return 1
torch.cuda = _Cuda()
torch.device = lambda *a: None
sys.modules['torch'] = torch
import mrhbm_loader
mrhbm_loader.load()
from lua_mapreduce_b200 import mrhbm
class Info:
    groups=1000; pairs_recv=1000; pairs_in=1000; sorted=1; key_bytes=8
class FakeMap:
    def __init__(s, c): s.c=c
    def gen_u64(s,*a): pass
    def gen_zipf(s,*a): pass
    def emit_batch_ptr(s,p,n): pass
    def wordcount(s,p,n=None): return 0
    def commit(s): pass
class FakeCtx:
    def __init__(s,*a,**k): s.num_partitions=a[1]
    def map_begin(s,j): return FakeMap(s)
    def shuffle(s): pass
    def stats(s): return dict(launches=5, ms_total=2.8, ms_combine=0.5, ms_hist=0.0, ms_plan=0.7, ms_scatter=0.7, ms_exchange=0.0, ms_sort_reduce=1.3, ms_bigbins=0.0, bins=56320, sub_bins=55, big_bins=0, attempts=1, bytes_exchanged=0)
    def result_info(s): return Info()
    def checksum_input(s): return [1,2,3,1000]
    def checksum_result(s): return [1,2,3,1000,0,0]
    def reset(s): pass
    def pinned_array(s,n,dt): return np.zeros(n, dtype=dt)
    def result_copy(s,k=None,v=None): return k,v,None
    def close(s): pass
mrhbm.Ctx = FakeCtx
import bench
bench.ClockSampler = type('CS',(),{'__init__':lambda s,d:None,'start':lambda s:None,'stop':lambda s:{'sm_mhz':1965.0,'sm_max_mhz':1965.0,'reasons':[],'samples':1}})
bench.wordcount_config1 = lambda job, a, table: {"value": 1.0}
sys.argv=['bench.py','--steps','2','--warmup','3','--pairs','1000','--cpu-sample','20000','--e2e-steps','1','--no-parity']
bench.main()
