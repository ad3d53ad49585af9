"""Output containers of the model plugins -- same surface as reference models/output_storage.py:8-126
(``VAEOutput.mods[mod].<field>``, ``set_with_dict``, ``unpack_values``), written for this package."""
import torch.distributions as _td

FIELDS = ("encoder_dist", "joint_dist", "joint_decoder_dist", "decoder_dist", "dec_dist_private", "latent_samples",
          "enc_dist_private", "cross_decoder_dist")
_DICT_FIELDS = ("latent_samples", "cross_decoder_dist")
fields = list(FIELDS)  # reference module attribute name


class ModalityOutput:
    """Per-modality record; distribution-valued fields must hold torch.distributions objects
    (reference output_storage.py:49-52), the two dict-valued ones must hold dicts."""
    __slots__ = ("id",) + FIELDS

    def __init__(self, id: str):
        self.id = id
        for f in FIELDS:
            setattr(self, f, None)

    @staticmethod
    def check_field_valid(field: str):
        assert field in FIELDS, "Unsupported field name {}. Choose out of: {}".format(field, list(FIELDS))

    @staticmethod
    def check_is_distribution(val, field):
        assert isinstance(val, _td.Distribution), \
            "{} value must be an instance of torch.distributions! Got: {}".format(field, val)

    def set_value(self, field: str, val):
        if val is not None:
            self.check_field_valid(field)
            if field in _DICT_FIELDS:
                assert isinstance(val, dict), "Expected {} to be a dict! Got {}".format(field, val)
            else:
                self.check_is_distribution(val, field)
        setattr(self, field, val)

    def get_value(self, field: str):
        self.check_field_valid(field)
        return getattr(self, field)


class VAEOutput:
    def __init__(self):
        self.mods = {}

    def add_new_modality(self, name: str):
        self.mods[name] = ModalityOutput(name)

    def set_value(self, mod: str, field: str, val):
        if mod not in self.mods:
            self.add_new_modality(mod)
        self.mods[mod].set_value(field, val)

    def set_with_dict(self, d, field: str):
        if d is not None:
            for key, val in d.items():
                self.set_value(key, field, val)

    def set_to_all(self, field: str, val):
        for key in self.mods:
            self.set_value(key, field, val)

    def get_all_values(self, field):
        return [m.get_value(field) for m in self.mods.values()]

    def unpack_values(self):
        return {f: self.get_all_values(f) for f in FIELDS}
