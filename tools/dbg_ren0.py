import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import coffe_b200
from oracle import coffe_oracle as O
ref=O.RefCoffe(); ref.set_power_spectrum(O.load_pk())
ref.set_sep(np.array([10.,20,40,100,150])/0.67); ref.set_z_mean([0.5,1.0,1.5,2.0]); ref.set_deltaz([0.3]*4); ref.set_l([0,2,4]); ref.set_contributions(("den","rsd","len","d1","d2","g1","g2","g3","g4","g5"))
ref.lib.ref_integrals(ref.h_)
eng=coffe_b200.Engine(0); pk=O.load_pk(); chi,_=ref.renorm_table()
res=eng.integrals(pk,pk[0][0],pk[0][-1],[(0,0)],divergent=True,chi_max=chi[-1])
x0,y0=ref.renorm0_table()
for i in range(0,40): print(i, x0[i], res["r40"][i], y0[i], res["renorm0"][i])
print(eng.last_timing())
