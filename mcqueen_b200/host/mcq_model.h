/* mcq_model.h - constants of the McQueen model (values: the reference's constants.c). */
#ifndef MCQ_MODEL_H_
#define MCQ_MODEL_H_

enum { MCQ_PRIOR_LOG_NORMAL = 0, MCQ_PRIOR_NORMAL = 1, MCQ_PRIOR_GAMMA = 2, MCQ_PRIOR_FLAT = 3,
       MCQ_PRIOR_EXPONENTIAL = 4 };                       /* constants.h:31-35 */

#define MCQ_INIT_MU 20.0              /* constants.c:17 */
#define MCQ_INIT_R 0.01               /* constants.c:20 */
#define MCQ_P_REC 0.01                /* constants.c:29 */
#define MCQ_LAMBDA_SEL 0.5            /* constants.c:33 */
#define MCQ_OMEGA_MU_SEL 1.0          /* constants.c:36 */
#define MCQ_OMEGA_SIGMA_SEL 2.0       /* constants.c:38 */
#define MCQ_OMEGA_GAMMA_K 1.0         /* constants.c:42 */
#define MCQ_OMEGA_GAMMA_THETA 20.0    /* constants.c:44 */
#define MCQ_MU_SEL_PRIOR 1.0          /* constants.c:48-49 (esc and rev) */
#define MCQ_SIGMA_SEL_PRIOR 2.0       /* constants.c:51-52 */
#define MCQ_MU_SEL_NORM_PRIOR 1.0     /* constants.c:56-57 */
#define MCQ_SIGMA_SEL_NORM_PRIOR 3.0  /* constants.c:59-60 */
#define MCQ_GAMMA_K_SHAPE 1.0         /* constants.c:64-65 */
#define MCQ_GAMMA_THETA_SCALE 20.0    /* constants.c:67-68 */
#define MCQ_INIT_WINDOWS 6            /* constants.c:71-72 */
#define MCQ_WINDOW_GEOM_PARAM 0.5     /* constants.c:75 */
#define MCQ_PRIOR_ADD_WINDOW_ESC 0.2  /* constants.c:78 */
#define MCQ_PRIOR_ADD_WINDOW_REV 0.3  /* constants.c:79 */
#define MCQ_MU_PRIOR 0.01             /* constants.c:82 */
#define MCQ_KAPPA 7.595744            /* constants.c:85 */
#define MCQ_PHI 0.9482171             /* constants.c:88 */

/* codon usage prior_C1 (constants.c:93-109) */
static const double mcq_prior_c1[64] = {
  0.01660719, 0.005966229, 0.04147626, 0.02127574, 0.0005901346, 0.0005592375, 0.0110179, 0.000117409,
  0.01344333, 0.0002224591, 1.235884e-05, 4.943536e-05, 0.008641918, 0.0005159815, 3.08971e-06, 0.02271246,
  0.004325594, 0.0001143193, 0.003095889, 0.005067124, 0.0264232, 0.01026711, 0.04408398, 0.000692095,
  0.01278522, 0.004915728, 0.04811605, 0.01067495, 6.179419e-06, 1.235884e-05, 0.0003027916, 0.00414639,
  0.01751247, 0.009772752, 0.03740403, 0.05353231, 0.0158224, 0.02728214, 0.02607715, 0.0002286385,
  0.03277255, 0.01464831, 0.03445644, 0.01345569, 0.01164512, 0.01513031, 0.03834948, 0.008382382,
  0.004328683, 0.009099195, 0.03378907, 0.01560921, 0.02094514, 0.01382336, 0.04780708, 0.005224699,
  0.02168976, 0.01999351, 0.03876968, 0.01819221, 0.002045388, 0.006686132, 0.04171726, 0.02556735};

#endif
