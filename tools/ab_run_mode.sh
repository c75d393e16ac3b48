for rep in 1 2 3; do for v in "" _run0 _run1; do TAROK_B200_LIB=$PWD/tarok_b200/libtarok_b200$v.so python - <<'PY'
import os,sys,torch,json
sys.path.insert(0,os.getcwd())
from tarok_b200.env import TarokEnv
n=1<<20
env=TarokEnv(n,seed=1); env.set_materialise(False); env.set_graph(False)
def steps():
    env.setup_synth(16,0)
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record(); env.step_random(48); b.record(); return a,b
steps(); torch.cuda.synchronize()
ev=[steps() for _ in range(20)]; torch.cuda.synchronize()
us=sum(a.elapsed_time(b) for a,b in ev)/len(ev)/48*1e3
a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
env.rollout(16,0); torch.cuda.synchronize(); a.record()
for i in range(20): env.rollout(16,i*n)
b.record(); torch.cuda.synchronize()
print(os.path.basename(os.environ["TAROK_B200_LIB"]).ljust(26),"step_random %.2f us  plain rollout %.1f us"%(us,a.elapsed_time(b)/20*1e3))
PY
done; done
