"""Model: an ordered dict of layers — drop-in for the reference's nn/model.py."""
from __future__ import annotations

from .. import distributed as npdist
from .core import Layer


class Model(Layer):
    def __init__(self) -> None:
        super().__init__()
        self.layers = {}

    def forward(self, inputs):
        for layer in self.layers.values():
            inputs = layer(inputs)
        return inputs

    def backward(self, d_out):
        for layer in list(self.layers.values())[::-1]:
            d_out = layer.backward(d_out)
        return d_out

    def add_layer(self, name, layer: Layer) -> None:
        setattr(self, name, layer)
        self.layers[name] = layer

    def get_layers(self):
        return {name: layer.get_layers() if isinstance(layer, Model) else layer
                for name, layer in self.layers.items()}

    def scatter(self, src: int = 0) -> None:
        assert not self.is_scattered, "Cannot scatter an already scattered model."
        for layer in self.layers.values():
            layer.scatter(src)
        self.is_scattered = True

    def all_gather(self) -> None:
        for layer in self.layers.values():
            layer.all_gather()
        self.is_scattered = False

    def to_dict(self) -> dict:
        layers = {name: layer.to_dict() for name, layer in self.layers.items()}
        return {"settings": self.settings, "layers": layers}

    @classmethod
    def from_dict(cls, model_dict: dict[str, dict]):
        settings, layers = model_dict["settings"], model_dict["layers"]
        settings |= {"weight_init": "zeros", "bias_init": "zeros"}
        model = cls(**settings)
        for name in model.layers:
            assert name in layers, f"Expected {name} in {layers}."
            model.add_layer(name, getattr(model, name).from_dict(layers[name]))
            if name == "mlp":  # parallel context must already be set up (reference: model.py:64-66)
                model.mlp.row_linear.use_bias = npdist.tensor_parallel_rank() == 0
        return model
