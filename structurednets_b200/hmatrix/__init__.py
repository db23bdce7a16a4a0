from structurednets_b200.hmatrix.hmatrix import (BlockClusterTree, HMatrix, HMatrixComponent, TreeElement,  # noqa: F401
                                                  approximate_hmatrix, build_hmat_block_cluster_tree)
