"""``FastAdversarialMF`` surface (FastAdversarialMF.py:13-145) on the APR engine.

PARITY UNPINNED, by design: the reference class is a popularity-discriminator GAN built on the un-vendored
third-party ``keras_adversarial`` (FastAdversarialMF.py:8-10,64-74; unpinned, absent from this image) with Adam and
an MSE head -- not APR.  BASELINE.json names the file for its Recommender-shaped surface, so this class keeps the
constructor, ``init`` and the six Recommender methods and trains with the adversarial (APR) step instead; the GAN is
not restated.  ``weight`` maps to reg_adv; ``pop_percent`` only drives the popular/rare split that ``init`` exposes.
"""
import numpy as np

from .MFRecommender import BPRRecommender


class FastAdversarialMF(BPRRecommender):
    def __init__(self, uNum, iNum, dim, weight, pop_percent, lr=0.05, eps=0.5, seed=2019):
        BPRRecommender.__init__(self, uNum, iNum, dim, lr=lr, adver=1, eps=eps, reg_adv=weight, seed=seed)
        self.weight = weight
        self.pop_percent = pop_percent

    def init(self, users, items):
        self.popular_user_x, self.rare_user_x = self.get_discriminator_train_data(users)
        self.popular_item_x, self.rare_item_x = self.get_discriminator_train_data(items)

    def get_discriminator_train_data(self, x):
        """FastAdversarialMF.py:129-145: ids by descending frequency, split at pop_percent."""
        ids, counts = np.unique(np.asarray(x), return_counts=True)
        order = np.argsort(-counts, kind="stable")
        ranked = ids[order]
        cut = int(len(ranked) * self.pop_percent)
        return ranked[:cut], ranked[cut:]

    def get_params(self):
        return "_w%g_pp%g%s" % (self.weight, self.pop_percent, BPRRecommender.get_params(self))
