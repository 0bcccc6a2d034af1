from udifftext_b200.host.conditioner import AbstractEmbModel, GeneralConditioner, LabelEncoder, LatentEncoder, SpatialRescaler  # noqa: F401
