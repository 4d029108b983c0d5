"""Priority strategies (mirror of train/prioritized_replay/generate_priority.py:4-58).  The reference driver only uses
LossPriority (train/__main__.py:142); TrendPriority / HybridPriority are unused there and broken under numpy 2 (`np.float`)
-- here they are restated with the same arithmetic, pinned by tests/golden/priority_strategies.npz."""
import numpy as np


class GeneratePriority:
    def get_priorities(self, batch_nodes_seed, losses):
        raise NotImplementedError


class LossPriority(GeneratePriority):
    """priority = per-vertex cross-entropy loss"""

    def get_priorities(self, batch_nodes_seed, losses):
        return losses


class TrendPriority(GeneratePriority):
    """priority = exponentially smoothed positive part of the loss increase of a vertex since it was last trained; a vertex
    seen for the first time starts from the running mean of all tracked values (generate_priority.py:11-46)"""

    def __init__(self, n_vertices, alpha=0.85):
        self.values = np.zeros(n_vertices, dtype=np.float64)
        self.prev_loss = np.zeros(n_vertices, dtype=np.float64)
        self.init = np.full(n_vertices, True, dtype=bool)
        self.avg = 0
        self.n_items = 0
        self.alpha = alpha

    def get_priorities(self, batch_nodes_seed, losses):
        idx = np.asarray(batch_nodes_seed)
        losses = np.asarray(losses, dtype=np.float64)
        fresh = idx[self.init[idx]]
        self.init[fresh] = False
        self.values[fresh] = self.avg
        self.n_items += len(fresh)
        gain = (losses - self.prev_loss[idx]).clip(min=0)
        # the running mean is maintained incrementally: remove the batch's old values, add the new ones
        self.avg *= self.n_items
        self.avg -= np.sum(self.values[idx])
        self.values[idx] *= self.alpha
        self.values[idx] += gain * (1 - self.alpha)
        self.avg += np.sum(self.values[idx])
        self.avg /= self.n_items
        self.prev_loss[idx] = losses
        return self.values[idx]


class HybridPriority(GeneratePriority):
    """(1 - loss_contrib) * trend + loss_contrib * loss (generate_priority.py:49-58)"""

    def __init__(self, n_vertices, alpha=0.85, loss_contrib=0.5):
        self.trend_p = TrendPriority(n_vertices, alpha)
        self.loss_p = LossPriority()
        self.loss_contrib = loss_contrib

    def get_priorities(self, batch_nodes_seed, losses):
        prior = self.trend_p.get_priorities(batch_nodes_seed, losses) * (1 - self.loss_contrib)
        prior += np.asarray(self.loss_p.get_priorities(batch_nodes_seed, losses), dtype=np.float64) * self.loss_contrib
        return prior
