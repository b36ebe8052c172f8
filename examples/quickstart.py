"""Quick start: the GerryChain ReCom recipe, unchanged, on the B200 engine - then the same
chain 1024 times at once.

    python examples/quickstart.py          (needs a CUDA device and the built library:
                                            python -c "import __graft_entry__ as g; g.build()")
"""

import functools
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import networkx
import numpy as np

from gerrychain_b200 import (BatchedMarkovChain, MarkovChain, Partition, Tally, always_accept, cut_edges, recom,
                             within_percent_of_ideal_population)
from gerrychain_b200.seed import random_plans

random.seed(2024)

# a dual graph: 40 x 40 lattice, random integer populations, 8 districts
graph = networkx.grid_2d_graph(40, 40)
for node in graph.nodes:
    graph.nodes[node]["TOTPOP"] = random.randint(50, 150)
k, eps = 8, 0.05

# seed plan: recursive_tree_part on the device (Partition.from_random_assignment)
initial = Partition.from_random_assignment(graph, k, eps, "TOTPOP",
                                           updaters={"population": Tally("TOTPOP", alias="population"),
                                                     "cut_edges": cut_edges})
ideal = sum(initial["population"].values()) / k
proposal = functools.partial(recom, pop_col="TOTPOP", pop_target=ideal, epsilon=eps, node_repeats=1)
constraints = [within_percent_of_ideal_population(initial, eps)]

# 1) the reference's driver loop, one chain
chain = MarkovChain(proposal, constraints, always_accept, initial, total_steps=200)
cuts = [len(state["cut_edges"]) for state in chain]
print("one chain, 200 steps: cut edges min/mean/max =", min(cuts), sum(cuts) / len(cuts), max(cuts))

# 2) 1024 chains resident on the GPU, each from its own random seed plan
plans = random_plans(graph, k, eps, "TOTPOP", n_plans=1024, seed=7)
batch = BatchedMarkovChain(proposal, constraints, always_accept, initial, total_steps=1001, n_chains=1024,
                           seed=11, initial_assignments=plans)
last = batch.run_to_end()
c = last.cut_edge_counts()
dev = np.abs(last.population() - ideal).max(axis=1) / ideal
print("1024 chains x 1000 steps: cut edges mean %.1f (sd %.1f), worst population deviation %.3f" % (c.mean(), c.std(), dev.max()))
print("plan of chain 0 as a Partition:", last.partition(0), "with", len(last.partition(0)["cut_edges"]), "cut edges")
