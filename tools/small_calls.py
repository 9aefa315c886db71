"""The reference's own bench shapes through the host API beside the single-thread CPU port (bench.py extras)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rulinalg_b200 as rla
import bench
l = rla.lib(); rla.check(l.rla_init(0))
print(json.dumps(bench.reference_shapes_extra(rla, l, np), indent=1))
