import sys, numpy as np
H=int(sys.argv[2]); a=np.fromfile(sys.argv[1],dtype=np.uint64).reshape(H,4,8).astype(np.int64)
t0=a[a>0].min()
names=['V','A','C','P']
lo,hi=H//2,H//2+400
for r,(nm,npts) in enumerate(zip(names,[6,6,6,3])):
    x=a[lo:hi,r,:npts]-t0
    row=np.diff(x[:,0])
    print(nm,'row period: mean %.0f med %.0f p90 %.0f'%(row.mean(),np.median(row),np.percentile(row,90)))
    seg=np.diff(x,axis=1)
    print('   segment means:',' '.join('%d->%d: %.0f'%(i,i+1,seg[:,i].mean()) for i in range(npts-1)), ' tail->next top: %.0f'%((x[1:,0]-x[:-1,npts-1]).mean()))
# lag between roles at row completion
x=a[lo:hi]-t0
print('lag V.arrive->A.slotwait done', (x[:,1,3]-x[:,0,5]).mean(), ' A.arrive->C.slot done',(x[:,2,3]-x[:,1,4]).mean(), ' C.arrive->V.freeP(row+K)?')
for K in (1,2,3,4): print('  K=%d: V.t4(row+K)-C.t4(row): %.0f'%(K,(x[K:,0,4]-x[:-K,2,4]).mean()))
print("total cycles/row", (a[hi-1,2,5]-a[lo,2,5])/(hi-1-lo))
