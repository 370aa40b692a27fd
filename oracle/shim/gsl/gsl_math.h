#ifndef MLMCPI_ORACLE_SHIM_GSL_MATH_H
#define MLMCPI_ORACLE_SHIM_GSL_MATH_H
struct gsl_function_struct {
  double (*function)(double x, void *params);
  void *params;
};
typedef struct gsl_function_struct gsl_function;
#define GSL_FN_EVAL(F, x) (*((F)->function))(x, (F)->params)
#endif
