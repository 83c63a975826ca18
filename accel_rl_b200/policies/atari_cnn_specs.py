"""CNN presets for 104x80 observations (reference: accel_rl/policies/atari_cnn_specs.py).
Presets 0 and 1 are supported by the CUDA path (one hidden layer, 8x8/4 first conv); 2-4 are listed
for completeness and rejected by the engine with a clear error."""

cnn_specs = dict()

cnn_specs["0"] = cnn_specs[0] = dict(   # standard "small": 900k params
    conv_filter_sizes=[8, 4], conv_filters=[16, 32], conv_strides=[4, 2], conv_pads=[(0, 0), (1, 1)],
    hidden_sizes=[256])
cnn_specs["1"] = cnn_specs[1] = dict(   # standard NIPS "large" (Nature-CNN padded for 104x80): 3.6M params
    conv_filter_sizes=[8, 4, 3], conv_filters=[32, 64, 64], conv_strides=[4, 2, 1],
    conv_pads=[(0, 0), (1, 1), (1, 1)], hidden_sizes=[512])
cnn_specs["2"] = cnn_specs[2] = dict(
    conv_filter_sizes=[5, 3, 3, 3, 3], conv_filters=[32, 64, 64, 128, 128], conv_strides=[3, 1, 1, 2, 1],
    conv_pads=[(0, 0), (1, 1), (1, 1), (1, 1), (1, 1)], hidden_sizes=[64, 64])
cnn_specs["3"] = cnn_specs[3] = dict(
    conv_filter_sizes=[4, 3, 3, 3, 3], conv_filters=[32, 64, 64, 64, 128], conv_strides=[2, 1, 1, 1, 2],
    conv_pads=[(0, 0), (1, 1), (1, 1), (1, 1), (0, 0)], hidden_sizes=[64, 64])
cnn_specs["4"] = cnn_specs[4] = dict(
    conv_filter_sizes=[16, 8, 4], conv_filters=[16, 32, 64], conv_strides=[3, 2, 1],
    conv_pads=[(1, 1), (1, 2), (1, 1)], hidden_sizes=[256])
