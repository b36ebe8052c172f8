"""Steps/s of the literal drop-in: MarkovChain + recom, one chain, host Partition objects
(the reference runs this loop at ~88 steps/s on Grid 20x20 and ~7 steps/s on Grid 100x100)."""
import functools, os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gerrychain_b200 import Grid, MarkovChain, recom
from gerrychain_b200.accept import always_accept
from gerrychain_b200.constraints import within_percent_of_ideal_population
from gerrychain_b200 import synthetic

for dims, k, steps in (((20, 20), 4, 2000), ((100, 100), 10, 1000)):
    random.seed(1)
    plan = None if k == 4 else synthetic.grid_stripes(dims[0], dims[1], k)
    grid = Grid(dims, assignment=plan)
    ideal = sum(grid["population"].values()) / len(grid)
    p = functools.partial(recom, pop_col="population", pop_target=ideal, epsilon=0.05)
    chain = MarkovChain(p, [within_percent_of_ideal_population(grid, 0.05)], always_accept, grid, steps)
    it = iter(chain); next(it); next(it)
    t0 = time.perf_counter(); n = 0
    for state in it:
        n += 1; _ = state.n_cut_edges
    dt = time.perf_counter() - t0
    print(f"Grid {dims[0]}x{dims[1]}, {k} districts: {n / dt:.0f} steps/s through MarkovChain/recom (n_cut_edges read every step)")
    t0 = time.perf_counter(); m = 0
    for state in MarkovChain(p, [within_percent_of_ideal_population(grid, 0.05)], always_accept, grid, 200):
        m += 1; _ = len(state["cut_edges"]); _ = state["population"]
    print(f"   with the cut_edges SET and the population dict materialised every step: {m / (time.perf_counter() - t0):.0f} steps/s")
