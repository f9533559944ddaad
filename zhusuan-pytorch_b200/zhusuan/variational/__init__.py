from .elbo import ELBO, EvidenceLowerBoundObjective
from .importance_weighted_objective import ImportanceWeightedObjective

# Route a Bernoulli likelihood under ImportanceWeightedObjective through the fused resident-column
# kernel (zs_iw_bernoulli_fused).  Set to False to force the general two-pass kernels, e.g. when the
# loss must be back-propagated more than once (retain_graph).
FUSED = True
