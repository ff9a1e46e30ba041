"""ModelFactory mirror (reference models/model_factory.py:14-186): the name -> class boundary of the hot path."""
import torch

from .modules import (UNet_AdditionFusion, UNet_AttentionOnly, UNet_Baseline, UNet_BothStandardConv, UNet_ConcatFusion,
                      UNet_DecoderOnlyDFC, UNet_EncoderOnlyDFC, UNet_FullResAttention, UNetDFCSARes)

# names the reference factory knows (models/model_factory.py:94-183) that are outside the B200 hot path
_OUT_OF_SCOPE = {
    "UNet", "TransformerUNet", "TransUNet", "ViTSegmentation", "ViT_Seg", "VisionTransformer",
}


class ModelFactory:
    def __init__(self, config=None):
        self.config = config

    def create_model(self, config=None):
        if config is None:
            if self.config is None:
                raise ValueError("必須提供配置")
            config = self.config
        return ModelFactory._create_model_impl(config)

    @staticmethod
    def get_model(config):
        """reference models/model_factory.py:51-72 (incl. the optional pretrained_path load that only prints on failure)."""
        model = ModelFactory._create_model_impl(config)
        if config["model"].get("pretrained_path"):
            try:
                model.load_state_dict(torch.load(config["model"]["pretrained_path"], weights_only=False))
                print(f"成功載入預訓練權重: {config['model']['pretrained_path']}")
            except Exception as e:  # noqa: BLE001 - same behaviour as the reference
                print(f"載入預訓練權重失敗: {e}")
        return model

    @staticmethod
    def _create_model_impl(config):
        m = config["model"]
        name = m["name"]
        in_channels = m.get("in_channels", 3)
        out_channels = m.get("out_channels", 1)
        features = m.get("features", [64, 128, 256, 512])
        pool_size = m.get("pool_size", 8)
        qk = m.get("ablation_on_qk_channels", 8)
        if name == "DFC-SA-Res-Block":                       # reference :103-110
            return UNetDFCSARes(in_channels=in_channels, out_channels=out_channels, features=features,
                                pool_size=pool_size, ablation_on_qk_channels=qk)
        if name == "UNet_FullResAttention":                  # reference :174-175 (ablation 3)
            return UNet_FullResAttention(in_channels=in_channels, out_channels=out_channels, features=features)
        # ablations 1, 2 and 4 (reference :162-171, :178-183): the same kernels re-wired
        if name == "UNet_AttentionOnly":
            return UNet_AttentionOnly(in_channels, out_channels, features, pool_size)
        if name == "UNet_AdditionFusion":
            return UNet_AdditionFusion(in_channels, out_channels, features, pool_size)
        if name == "UNet_ConcatFusion":
            return UNet_ConcatFusion(in_channels, out_channels, features, pool_size)
        if name == "UNet_Baseline":
            return UNet_Baseline(in_channels, out_channels, features)
        if name == "UNet_BothStandardConv":
            return UNet_BothStandardConv(in_channels, out_channels, features)
        if name == "UNet_EncoderOnlyDFC":
            return UNet_EncoderOnlyDFC(in_channels, out_channels, features, pool_size)
        if name == "UNet_DecoderOnlyDFC":
            return UNet_DecoderOnlyDFC(in_channels, out_channels, features, pool_size)
        if name in _OUT_OF_SCOPE:
            raise NotImplementedError(f"dfcsa: model '{name}' is outside the DFC-SA-Res-Block hot path this library accelerates")
        raise ValueError(f"不支援的模型類型: {name}")   # reference models/model_factory.py:186
