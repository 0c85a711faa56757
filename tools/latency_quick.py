import sys, os, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import ptts_b200 as P
from make_assets import default_model_dir
d = default_model_dir(eos_mode="never")
ctx = P.Context(d, max_slots=1, kv_capacity=1024)
P.set_seed(0)
st = ctx.stream("cosette", temp=0.0)
SENT = "The quick brown fox jumped over the sleeping dog."
for rep in range(4):
    st.reset()
    t0 = time.perf_counter(); st.send(SENT); t1 = time.perf_counter(); st.flush(); t2 = time.perf_counter()
    f = st.receive(); t3 = time.perf_counter()
    f = st.receive(); t4 = time.perf_counter()
    n = 2
    while st.receive() is not None: n += 1
    t5 = time.perf_counter()
    print(f"LAT rep{rep} send {1e3*(t1-t0):.3f} flush {1e3*(t2-t1):.3f} first_receive {1e3*(t3-t2):.3f} second {1e3*(t4-t3):.3f} rest/frame {1e3*(t5-t4)/(n-2):.3f} ms")
