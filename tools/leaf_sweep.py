"""SAH leaf-size / cost sweep for the 871 200-triangle mesh room (RTB200_MAX_LEAF, RTB200_COST_PRIM are read at commit)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ab_pool
from ray_tracing_series_rust_b200 import capi
for leaf, cost in (("4", "2.0"), ("8", "2.0"), ("8", "1.0"), ("8", "0.5"), ("6", "1.0"), ("4", "1.0"), ("2", "2.0")):
    os.environ["RTB200_MAX_LEAF"] = leaf; os.environ["RTB200_COST_PRIM"] = cost
    s = ab_pool.scene(14, 0xB004, 660)
    s.render(capi.make_config(1000, 1.0, 4, 50))
    hc = s.host_check()
    ab_pool.timed(s, 1000, 1.0, 20, f"mesh871k max_leaf {leaf} cost_prim {cost} nodes {hc.get('nodes')} leaves {hc.get('leaves')}", 0)
    s.close()
