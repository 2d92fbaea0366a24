import sys
import numpy as np
t=np.load(sys.argv[1])
n=len(t); t0=t[:,3].min()
pub=(t[:,0]-t0)/1e3; res=(t[:,1]-t0)/1e3; tick=(t[:,3]-t0)/1e3
spins=t[:,2]&0xffffffff; last=(t[:,2]>>32)&1
ns=(n+31)//32
comp=np.array([pub[j*32:(j+1)*32].max() for j in range(ns)])
first=np.array([pub[j*32:(j+1)*32].min() for j in range(ns)])
la=np.where(last==1)[0]
inc=np.full(ns,np.nan); inc[la//32]=res[la]
print("super: spread of publishes within a super (comp-first): median %.2f p90 %.2f"%(np.median(comp-first),np.percentile(comp-first,90)))
print("super: inc-comp median %.2f p90 %.2f"%(np.nanmedian(inc-comp), np.nanpercentile(inc-comp,90)))
cm=np.maximum.accumulate(comp)
print("super: inc - cummax(comp) (pure look-back latency once all earlier are complete): median %.2f p90 %.2f"%(np.nanmedian(inc-cm),np.nanpercentile(inc-cm,90)))
j=np.arange(n)//32
prev_inc=np.where(j>0, inc[np.maximum(j-1,0)], 0)
d=res-np.maximum(prev_inc,pub)
print("tile resolved - max(INC(j-1), own publish): median %.2f p90 %.2f"%(np.median(d),np.percentile(d,90)))
lag=np.maximum.accumulate(pub)-pub
print("publish out-of-order lag (cummax(pub)-pub): median %.2f p90 %.2f max %.2f"%(np.median(lag),np.percentile(lag,90),lag.max()))
print("us per spin: median %.2f"%np.median((res-pub)/np.maximum(spins,1)))
print("ticket->publish median %.2f ; publish->resolved median %.2f ; resolved->next ticket?"%(np.median(pub-tick),np.median(res-pub)))
