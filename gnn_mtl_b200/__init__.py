"""eagraft — the GNN-MTL entity-alignment hot path on B200 (sm_100a).

Python mirror of the reference's operator API over the C ABI in include/eagraft.h:

    gnn_mtl_b200.layers.layers        GraphConvolution, HighWayGraphConvolution, Linear, get_dim_act
    gnn_mtl_b200.models               encoders / decoders / models_ea (EAModel, UEAModel, BaseModel.get_neg)
    gnn_mtl_b200.utils.ot_loss        sinkhorn (drop-in), sinkhorn_fused (cost never materialised)
    gnn_mtl_b200.SinkhornOT           sinkhorn_iteration, cost matrices
    gnn_mtl_b200.utils.eval_utils     get_hits, eval_at_1, eval_gw_matching_matrix
    gnn_mtl_b200.utils.data_utils     get_sparse_tensor, DBP15K loaders
    gnn_mtl_b200.parallel             one-node multi-GPU (row-sharded Sinkhorn / eval / SpMM)
    gnn_mtl_b200.run                  the two training schedules

Importing any kernel-backed module loads gnn_mtl_b200/csrc/libeagraft.so and fails loudly without it.
"""
__version__ = "0.1.0"
