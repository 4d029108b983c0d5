"""Backend-agnostic trainer base (mirror of train/graphsage/model.py:18-117): train_timestep timing,
evaluation glue, macro-F1 + confusion matrix appended to the result CSV."""
import time

import numpy as np
import sklearn.metrics
import torch
from sklearn.metrics import f1_score


class SupervisedGraphSage:
    def __init__(self, graphsage_model, batch_per_timestep, batch_size, labels, samples, n_workers, cuda, batch_full):
        self.graphsage_model = graphsage_model
        self.batch_size = batch_size
        self.batch_per_timestep = batch_per_timestep
        self.samples = samples
        self.n_workers = n_workers
        self.labels = labels
        self._cuda_var = cuda
        self.batch_full = batch_full
        self.amount_of_train = {}
        self.delay = 0.0

    def build_optimizer(self):
        raise NotImplementedError

    def get_model(self):
        return "base_model"

    def choose_vertices(self, graph_util):
        raise NotImplementedError

    def _run_custom_eval(self, graph, subgraph_to_id, id_to_subgraph, test_vertices):
        raise NotImplementedError

    def _run_custom_train(self, graph, subgraph_to_id, id_to_subgraph, train_vertices, graph_util):
        raise NotImplementedError

    # ---- training ----------------------------------------------------------------------------------
    def train_timestep(self, graph_util):
        """`delay` spans _run_custom_train only, like the reference (:108-117); the device is drained
        before the clock stops because every launch is asynchronous."""
        batch_nodes = self.choose_vertices(graph_util)
        start = time.time()
        id_to_subgraph = graph_util.get_original_to_subgraph_map()
        subgraph_to_id = graph_util.get_subgraph_to_original_map()
        graph = graph_util.get_graph()
        self._run_custom_train(graph, subgraph_to_id, id_to_subgraph, id_to_subgraph[batch_nodes], graph_util)
        torch.cuda.synchronize()
        self.delay = time.time() - start

    # ---- evaluation --------------------------------------------------------------------------------
    def evaluate(self, graph_util, path):
        return self._evaluate_vertices(graph_util, path, np.array(graph_util.get_test_set()))

    def evaluate_next_snapshots(self, temporal_graph, delta, path, at_least=20):
        new_vertices, labelled = temporal_graph.get_added_vertices(delta)
        test = np.array(new_vertices)[np.asarray(labelled, dtype=bool)]
        if len(test) < at_least:
            if path:
                with open(path, "a+") as f:
                    f.write(self.get_model() + ";;;\n")
            return None
        return self._evaluate_vertices(temporal_graph, path, test)

    def _evaluate_vertices(self, graph_util, path, batch_nids):
        id_to_subgraph = graph_util.get_original_to_subgraph_map()
        subgraph_to_id = graph_util.get_subgraph_to_original_map()
        graph = graph_util.get_graph()
        vertices = np.asarray(id_to_subgraph[batch_nids], dtype=np.int64)
        device_eval = getattr(self, "_eval_logits_device", None)
        if device_eval is not None and type(self)._run_custom_eval is getattr(type(self), "_base_run_custom_eval", None):
            # logits never leave the GPU: argmax + confusion matrix in one kernel (ogl_eval_confusion), C x C int64 come back and
            # the macro-F1 follows from the matrix exactly as sklearn derives it from (labels, predictions) (reference :83-86)
            from .._native import eval_confusion, macro_f1_from_confusion
            logits_dev = device_eval(graph, vertices)
            if logits_dev is None or logits_dev.shape[0] == 0:
                return None
            labels_dev = graph.ndata["target"][torch.as_tensor(vertices, device="cuda")].reshape(-1)
            cm_dev, _ = eval_confusion(logits_dev, labels_dev)
            f1, cm = macro_f1_from_confusion(cm_dev.cpu().numpy())
        else:
            # a subclass overrode the reference's hook (returns host chunks): the reference's own host path
            chunks = self._run_custom_eval(graph, subgraph_to_id, id_to_subgraph, vertices)
            if len(chunks) == 0:
                return None
            logits = np.concatenate(chunks)
            if len(logits) == 0:
                return None
            labels = graph.ndata["target"][torch.as_tensor(vertices, device="cuda")].reshape(-1).cpu().numpy()
            pred = logits.argmax(axis=1)
            cm = sklearn.metrics.confusion_matrix(labels, pred)
            f1 = f1_score(labels, pred, average="macro")
        if path:
            with open(path, "a+") as f:
                f.write(self.get_model() + ";" + str(f1) + ";" + str(self.delay) + ";" + str([int(x) for r in cm for x in r]) + "\n")
        return f1

    def generate_tsne(self, graph_util, folder, index):
        raise NotImplementedError("TSNE plotting is out of scope (the reference's only call site is commented out, "
                                  "train/__main__.py:188-189)")
