"""Host-side H-matrix data model (construction only; the per-batch work is in csrc/hmat.cu).

The layer only needs the flattened list of leaves ``(row_range, col_range, left_lr, right_lr)``
("hmatrix/ tree ... data model only").  This module provides an object with the
interface ``HMatLayer`` reads from the reference's ``HMatrix`` (hmatrix/hmatrix.py:7-70:
``clone()``, ``get_all_hmatrix_components()``, ``get_nb_params()``, ``to_dense()``, ``shape``) plus a
from-scratch builder so the layer's default constructor works: an eta-admissibility quad-tree with
the reference's split rule (hmatrix/tree_element.py:119-162, approximators/hmat_approximator.py:9-20)
and the greedy "best error reduction per added parameter" singular-value allocation
(hmatrix/hmatrix.py:41-64), implemented here with one SVD per leaf and a heap instead of repeated
re-evaluation.  Reference ``HMatrix`` objects are accepted as well (duck typing).
"""
import heapq
import pickle

import numpy as np
import torch
import torch.nn as nn


class HMatrixComponent(nn.Module):
    """One leaf: rows x cols block stored as ``left_lr (rows x k) @ right_lr (k x cols)``; parameter
    names match the reference (hmatrix/hmatrix_component.py:5-21)."""

    def __init__(self, row_range: range, col_range: range, left_lr, right_lr):
        super().__init__()
        self.row_range = row_range
        self.col_range = col_range
        if left_lr is not None and right_lr is not None:
            self.left_lr = nn.Parameter(left_lr.detach().clone().float())
            self.right_lr = nn.Parameter(right_lr.detach().clone().float())
        else:
            self.left_lr = None
            self.right_lr = None

    def are_low_rank_components_set(self) -> bool:
        return self.left_lr is not None and self.right_lr is not None

    def get_component_shape(self):
        return (len(self.row_range), len(self.col_range))

    def get_nb_params(self) -> int:
        return int(self.left_lr.numel() + self.right_lr.numel()) if self.are_low_rank_components_set() else 0

    def to_dense(self) -> torch.Tensor:
        if self.are_low_rank_components_set():
            return torch.matmul(self.left_lr, self.right_lr)
        return torch.zeros(self.get_component_shape())


class TreeElement:
    def __init__(self, children, row_range: range, col_range: range):
        self.children = children
        self.row_range = row_range
        self.col_range = col_range
        self.hmatrix_component = HMatrixComponent(row_range, col_range, None, None)

    def is_leaf(self) -> bool:
        return self.children is None

    def set_hmatrix_component(self, left_lr: torch.Tensor, right_lr: torch.Tensor):
        assert left_lr.shape[0] == len(self.row_range), "The size of the left low rank component should match the size of the row range"
        assert right_lr.shape[1] == len(self.col_range), "The size of the right low rank component should match the size of the column range"
        assert left_lr.shape[1] == right_lr.shape[0], "The shapes of the left and right low rank components should match such that they can be multiplied"
        self.hmatrix_component = HMatrixComponent(self.row_range, self.col_range, left_lr, right_lr)

    def get_row_col_dist(self) -> int:
        r, c = self.row_range, self.col_range
        if c.stop < r.start:
            return r.start - c.stop
        if (r.start <= c.stop <= r.stop) or (r.start <= c.start <= r.stop):
            return 0
        return c.start - r.stop

    def is_admissible(self, eta: float) -> bool:
        return max(len(self.row_range), len(self.col_range)) <= 2 * eta * self.get_row_col_dist()

    def leaves(self):
        if self.is_leaf():
            return [self]
        out = []
        for ch in self.children:
            out += ch.leaves()
        return out


class BlockClusterTree:
    def __init__(self, root: TreeElement):
        self.root = root

    def get_all_leaf_elements(self):
        return self.root.leaves()

    def get_all_leaf_ranges(self):
        return [(leaf.row_range, leaf.col_range) for leaf in self.root.leaves()]

    def get_all_hmatrix_components(self):
        return [leaf.hmatrix_component for leaf in self.root.leaves()]

    def get_nb_params(self) -> int:
        return sum(c.get_nb_params() for c in self.get_all_hmatrix_components())


def build_hmat_block_cluster_tree(matrix_shape: tuple, eta: float, min_block_size=2) -> BlockClusterTree:
    assert eta > 0, "Eta must be greater than 0"
    assert eta < 1, "Eta must be smaller than 1"
    assert min_block_size > 1, "The minimum block size must be greater than 1"
    root = TreeElement(None, range(matrix_shape[0]), range(matrix_shape[1]))
    stack = [root]
    while stack:
        el = stack.pop()
        r, c = el.row_range, el.col_range
        if len(r) > 2 * min_block_size and len(c) > 2 * min_block_size and not el.is_admissible(eta):
            rm = int((r.stop + r.start) / 2)
            cm = int((c.stop + c.start) / 2)
            el.children = [TreeElement(None, range(r.start, rm), range(c.start, cm)), TreeElement(None, range(rm, r.stop), range(c.start, cm)),
                           TreeElement(None, range(r.start, rm), range(cm, c.stop)), TreeElement(None, range(rm, r.stop), range(cm, c.stop))]
            stack.extend(el.children)
    return BlockClusterTree(root)


class HMatrix:
    def __init__(self, block_cluster_tree: BlockClusterTree, shape=None):
        self.block_cluster_tree = block_cluster_tree
        if shape is not None:
            self.shape = shape

    def get_all_hmatrix_components(self) -> list:
        return self.block_cluster_tree.get_all_hmatrix_components()

    def get_nb_params(self) -> int:
        return self.block_cluster_tree.get_nb_params()

    def to_dense(self) -> torch.Tensor:
        res = torch.zeros(self.shape)
        for c in self.get_all_hmatrix_components():
            res[c.row_range.start:c.row_range.stop, c.col_range.start:c.col_range.stop] = c.to_dense().detach()
        return res

    def to_dense_numpy(self) -> np.ndarray:
        return self.to_dense().detach().numpy()

    def clone(self):
        return pickle.loads(pickle.dumps(self))


def _leaf_svds(optim_mat: np.ndarray, leaves, device=None):
    """One (U, S, Vh) per leaf block.  On a CUDA device the leaves are grouped by shape and every group is ONE batched float64 SVD
    (cuSOLVER through torch.linalg.svd): the 2048 -> 1000 layer has 2 560 leaves, 1 574 of them 4 x 8."""
    blocks = [optim_mat[l.row_range.start:l.row_range.stop, l.col_range.start:l.col_range.stop] for l in leaves]
    if device is None:
        return [np.linalg.svd(b, full_matrices=False) for b in blocks]
    out = [None] * len(leaves)
    groups = {}
    for i, b in enumerate(blocks):
        groups.setdefault(b.shape, []).append(i)
    for shape, idx in groups.items():
        if min(shape) == 0:
            for i in idx:
                out[i] = (np.zeros((shape[0], 0)), np.zeros((0,)), np.zeros((0, shape[1])))
            continue
        batch = torch.as_tensor(np.stack([blocks[i] for i in idx]), dtype=torch.float64, device=device)
        U, S, Vh = torch.linalg.svd(batch, full_matrices=False)
        U, S, Vh = U.cpu().numpy(), S.cpu().numpy(), Vh.cpu().numpy()
        for j, i in enumerate(idx):
            out[i] = (U[j], S[j], Vh[j])
    return out


def approximate_hmatrix(optim_mat: np.ndarray, nb_params_share: float, eta: float = 0.5, min_block_size: int = 2, device=None) -> HMatrix:
    """Greedy singular-value allocation: while the budget allows, give one more singular value to the leaf
    whose squared-error reduction per added parameter, sigma_{k+1}^2 / (rows + cols), is largest; a leaf
    can grow while its rank is below both of its dimensions (reference hmatrix/hmatrix.py:41-64,
    hmatrix/hmatrix_component.py:54-59)."""
    shape = optim_mat.shape
    tree = build_hmat_block_cluster_tree(shape, eta=eta, min_block_size=min_block_size)
    budget = int(shape[0] * shape[1] * nb_params_share)
    leaves = tree.get_all_leaf_elements()
    svds, ranks, heap = _leaf_svds(np.asarray(optim_mat), leaves, device), [0] * len(leaves), []
    for i, leaf in enumerate(leaves):
        S = svds[i][1]
        cost = len(leaf.row_range) + len(leaf.col_range)
        if len(S) > 0:
            heapq.heappush(heap, (-(S[0] ** 2) / cost, i))
    used = 0
    while heap:
        _, i = heapq.heappop(heap)
        leaf = leaves[i]
        rows, cols = len(leaf.row_range), len(leaf.col_range)
        cost = rows + cols
        if used + cost > budget:
            continue     # this leaf can never be afforded again (budget only shrinks)
        ranks[i] += 1
        used += cost
        k = ranks[i]
        if k < rows and k < cols and k < len(svds[i][1]):
            heapq.heappush(heap, (-(svds[i][1][k] ** 2) / cost, i))
    for i, leaf in enumerate(leaves):
        k = ranks[i]
        if k > 0:
            U, S, Vh = svds[i]
            root_s = np.sqrt(S[:k])
            leaf.set_hmatrix_component(torch.tensor(U[:, :k] * root_s[None, :]).float(), torch.tensor(root_s[:, None] * Vh[:k, :]).float())
    return HMatrix(tree, shape=tuple(shape))
