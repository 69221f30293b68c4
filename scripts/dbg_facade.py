import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from self_play_reinforcement_learning_b200 import envs
from self_play_reinforcement_learning_b200.mcts import MCTreeSearch
p = MCTreeSearch(None, envs.Connect4Env, iterations=20, net="hash", seed=17)
print("after ctor", p._slot(), p.game_index)
p.evaluate(True)
p.reset(player=1)
print("after reset", p._slot(), p.game_index)
for i in range(6):
    p._engine.run_ticks(4)
    print(i, p._slot(), p._engine.counters())
