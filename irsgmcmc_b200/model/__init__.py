from .loss import GMM, SSD, DataLoss, Entropy, EntropyMultivariateNormal, RegLoss, RegLoss_L2, RegLoss_LogNormal
from .distributions import (DirichletPrior, LogEnergyExpGammaPrior, LogPrecisionExpGammaPrior, LogScaleNormalPrior,
                            NormalDistribution)
