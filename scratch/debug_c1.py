import sys; sys.path.insert(0,'.')
import numpy as np
from oracle import oracle as O
from tests import util
import ggcat_b200 as G
recs=util.c1_records(); reads=O.Reads.from_list(recs)
k,m,b1,b2,s=31,12,2,6,1
sk,_=O.bucketing(reads,k,m,b1,b2)
ctx,st=G.minimizer_bucketing([(reads.data,reads.offsets)],b1,b2,k,m,min_multiplicity=s)
nbad=0
for b in range(5):
    tab=ctx.merge_bucket_range(b,1)
    for u in range(b<<b2,(b+1)<<b2):
        ref,_,tk=O.merge_unit(reads,sk,u>>b2,u&63,k,s)
        ref=ref[ref['kept']==1]
        sl=tab.unit_slice(u)
        ok=np.array_equal(tab.keys_lo[sl],ref['key_lo']) and np.array_equal(tab.multiplicity[sl].astype(np.uint64),ref['multiplicity']) and np.array_equal(tab.flags[sl],ref['flags'])
        if not ok:
            nbad+=1
            if nbad<4:
                gk=tab.keys_lo[sl]; print('unit',u,'n_gpu',len(gk),'n_ref',len(ref),'tk',tk)
                if len(gk)==len(ref):
                    d=np.nonzero((gk!=ref['key_lo'])|(tab.multiplicity[sl]!=ref['multiplicity'])|(tab.flags[sl]!=ref['flags']))[0]
                    print(' diffs',len(d),d[:5]); 
                    for i in d[:5]: print('  ',hex(int(gk[i])),int(tab.multiplicity[sl][i]),int(tab.flags[sl][i]),'ref',hex(int(ref['key_lo'][i])),int(ref['multiplicity'][i]),int(ref['flags'][i]),int(ref['counter'][i]))
print('bad units',nbad)
