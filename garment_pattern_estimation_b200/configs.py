"""Configuration VALUES of the shipped models (no code): what the reference reads from its YAML files, for callers that
build the models without the reference tree (bench.py, tools/, examples).

  ATT_NN_CONFIG / ATT_DATA_CONFIG   models/att/att.yaml:44-51,89-120 (+ 23 panel classes => max_pattern_len 23,
                                    nn/data/datasets.py:377-379)
  BASELINE_NN_CONFIG                models/baseline/lstm_stitch_tags.yaml:96-122
  ATT_STANDARDIZE                   models/att/att.yaml:61-73 (ground-truth shift / scale used by the loss and its metrics)
"""

ATT_NN_CONFIG = {
    'model': 'GarmentSegmentPattern3D', 'feature_extractor': 'EdgeConvFeatures', 'conv_depth': 2,
    'k_neighbors': 5, 'EConv_hidden': 200, 'EConv_hidden_depth': 2, 'EConv_feature': 150, 'EConv_aggr': 'max',
    'global_pool': 'mean', 'skip_connections': True, 'graph_pooling': False, 'pool_ratio': 0.1,
    'local_attention': True, 'panel_decoder': 'LSTMDecoderModule', 'panel_encoding_size': 250,
    'panel_hidden_size': 250, 'panel_n_layers': 3, 'lstm_init': 'kaiming_normal_',
    'pattern_decoder': 'LSTMDecoderModule', 'pattern_encoding_size': 250, 'pattern_hidden_size': 250,
    'pattern_n_layers': 2, 'stitch_tag_dim': 3,
}

ATT_DATA_CONFIG = {
    'max_pattern_len': 23, 'max_panel_len': 14, 'element_size': 4, 'rotation_size': 4, 'translation_size': 3,
}

BASELINE_NN_CONFIG = {
    'model': 'GarmentFullPattern3D', 'feature_extractor': 'EdgeConvFeatures', 'conv_depth': 2, 'k_neighbors': 5,
    'EConv_hidden': 200, 'EConv_hidden_depth': 2, 'EConv_feature': 150, 'EConv_aggr': 'max', 'global_pool': 'mean',
    'skip_connections': False, 'graph_pooling': False, 'pool_ratio': 0.1, 'local_attention': True,
    'panel_decoder': 'LSTMDecoderModule', 'panel_encoding_size': 250, 'panel_hidden_size': 250, 'panel_n_layers': 3,
    'lstm_init': 'kaiming_normal_', 'pattern_decoder': 'LSTMDecoderModule', 'pattern_encoding_size': 250,
    'pattern_hidden_size': 250, 'pattern_n_layers': 2,
}

ATT_STANDARDIZE = {
    'gt_shift': {'outlines': [0, 0, 0.14890235662460327, 0.05642016604542732],
                 'rotations': [-0.7071067690849304, -0.9238795042037964, -1, 0],
                 'translations': [-55.255470275878906, -20.001333236694336, -17.086795806884766]},
    'gt_scale': {'outlines': [25.267892837524418, 31.298505783081055, 0.2677369713783264, 0.2352069765329361],
                 'rotations': [1.7071068286895752, 1.9238795042037964, 1.7071068286895752, 1],
                 'translations': [109.58930206298828, 98.27909088134766, 37.84679412841797]},
}
