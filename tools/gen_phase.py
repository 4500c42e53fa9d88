import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import lstm_gen_check as c
for (H, B, Tn) in ((1024, 64, 80), (1024, 128, 80)):
    print("case", H, B, Tn, flush=True)
    c.run(H, B, Tn, stops=False, time_it=True, bwd=(B <= 64))
