"""fastest-lap hot path, B200-native: collocation NLP callbacks on sm_100a."""
from .vehicle import Vehicle, vehicle_from_xml, MODEL_F1, MODEL_KART  # noqa: F401
from .track import Track, track_from_xml  # noqa: F401
