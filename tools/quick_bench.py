import sys, time
sys.path.insert(0,'/root/repo')
import miso_b200 as mb
for kind,G,R in ((0,2000,1000),(1,3000,2000)):
    t=time.time(); w=mb.Workload(kind,G,R,36,250.,900.,4.,seed=1); t1=time.time()
    plan=mb.Plan().append(w); t2=time.time()
    params=mb.make_params(5000,500,10,1,seed=1)
    plan.upload(params)
    for rep in range(2):
        ms,nl=plan.run_resident()
        print('kind',kind,'G',G,'gen %.2fs plan %.2fs kernel %.1f ms launches %d -> %.3g iters/s'%(t1-t,t2-t1,ms,nl,G*5000/(ms/1e3)),flush=True)
    out=plan.download()
    print(plan.gene_result(out,0)['samples'].mean(axis=1), w.truth(0, plan.info()[0,0]))
