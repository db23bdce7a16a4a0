"""structurednets_b200 -- B200-native (sm_100a) forward/backward of the structured-weight layers of
MatthiasKi/structurednets, behind the reference's own nn.Module constructors.  See DESIGN.md."""
__version__ = "0.1.0"
