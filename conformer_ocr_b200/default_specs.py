"""Architecture defaults of the reference (conformer_ocr/default_specs.py:48-61).  Only the keys the
hot path consumes are kept; optimiser / scheduler / augmentation settings are out of scope."""

RECOGNITION_HYPER_PARAMS = {
    'height': 96,
    'encoder_dim': 144,
    'num_encoder_layers': 16,
    'num_attention_heads': 4,
    'feed_forward_expansion_factor': 4,
    'conv_expansion_factor': 2,
    'input_dropout_p': 0.1,
    'feed_forward_dropout_p': 0.1,
    'attention_dropout_p': 0.1,
    'conv_dropout_p': 0.1,
    'conv_kernel_size': 31,
    'half_step_residual': True,
    'subsampling_conv_channels': 32,
    'subsampling_factor': 4,
}
