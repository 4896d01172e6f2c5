"""Host-side mirror of the reference's layer / builder interface for the CTR hot path
(same class names, constructor arguments and error behaviour), as torch.nn.Modules whose
forward/backward run the hand-written sm_100a kernels through the C-ABI.

    reference module                      ->  here
    InteractingLayer.py                   ->  api.interacting_layer.InteractingLayer
    rank/multi_head/interacting_layer.py  ->  (same class)
    din.py                                ->  api.din.DIN
    staytime/layer.py                     ->  api.staytime_layer.{DIN, DeepCrossLayer, FMLayer}
    rough_rank/layer.py                   ->  api.rough_rank_layer.{DNN, MMOE, PLE, CrossNet, Similarity, KDLoss}
    tn.feature_column / tn.layers / tn.core -> api.embedding.{category_column, embedding_column,
                                               EmbeddingFeatures, Adam, AdaGrad}
    autoint, rank/multi_head/multidnn.py  ->  api.builders.{AutoInt, AUTOINT}
    staytime/VideoDnn.py, model.py, config.py -> api.video_dnn.{create_moe_sub_model, mtl_net, create_model_func,
                                               custom_kl_loss, cross_entropy, mse_loss, huber_loss}, api.staytime_config
    rank/ctr/base_model.py, model_init.py ->  api.rank_ctr.{parse_feature_slots, RankCtrSubModel, Model}
    rough_rank/model.py, config/config.py ->  api.rough_rank_model.{create_tower, create_tower_teacher,
                                               create_shallow_tower, DSSM, create_model, mse_loss, y_pred_loss, config}
"""
from .interacting_layer import InteractingLayer  # noqa: F401
from .din import DIN  # noqa: F401
