"""Priority strategies (mirror of train/prioritized_replay/generate_priority.py:4-9; the Trend /
Hybrid strategies are unused by the reference driver and broken under numpy 2 -- SURVEY 8(f)-4)."""


class GeneratePriority:
    def get_priorities(self, batch_nodes_seed, losses):
        raise NotImplementedError


class LossPriority(GeneratePriority):
    """priority = per-vertex cross-entropy loss"""

    def get_priorities(self, batch_nodes_seed, losses):
        return losses
