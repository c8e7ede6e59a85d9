# round 2, very last call: warp_peers in four instructions per digit bit (R2P + VOTE + predicated NOT + OR) -- time and layout parity
timeout 75 python tools/ab_scatter.py peers4 check > gpurun_out/r02z_ab.log 2>&1; cat gpurun_out/r02z_ab.log
