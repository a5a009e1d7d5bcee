import sys, torch
sys.path.insert(0, ".")
from qppvm_b200 import api, gen
from qppvm_b200.layout import CONFIGS
d = CONFIGS[1]["desc"]; rob = gen.robot_for(d.n_a); s = api.Solver(d); s.set_robot(rob, (rob.foot + rob.hand)[:d.n_contacts])
st = torch.from_numpy(gen.generate_states(d, 2048, 1)).cuda().repeat(16, 1).contiguous()
rec = s.records_from_states(st)
for _ in range(3):
    s.records_from_states(st, records=rec)
torch.cuda.synchronize()
