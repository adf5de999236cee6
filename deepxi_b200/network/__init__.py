"""Mirror of deepxi/network: selector, tcn.ResNetV2, attention.MHANetV3."""
