from .conv_bias_relu import (ConvBias, ConvBias_, ConvBiasMaskReLU, ConvBiasMaskReLU_, ConvBiasReLU, ConvBiasReLU_, ConvFrozenScaleBiasReLU,
                             ConvFrozenScaleBiasReLU_)

__all__ = ["ConvBiasReLU", "ConvBiasMaskReLU", "ConvBias", "ConvFrozenScaleBiasReLU", "ConvBiasReLU_", "ConvBiasMaskReLU_", "ConvBias_",
           "ConvFrozenScaleBiasReLU_"]
