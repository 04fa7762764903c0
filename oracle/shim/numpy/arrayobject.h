/* numpy C-API stand-in for compiling the reference's nblist / forcefield sources as plain C (see ../Python.h). */
#ifndef EHIC_ORACLE_SHIM_ARRAYOBJECT_H
#define EHIC_ORACLE_SHIM_ARRAYOBJECT_H

#include <Python.h>

typedef long npy_intp;
typedef struct { PyObject_HEAD char *data; int nd; npy_intp *dimensions; npy_intp *strides; } PyArrayObject;

extern PyTypeObject PyArray_Type;
#define PyArray_Check(op) (((PyObject *)(op))->ob_type == &PyArray_Type)
#define PyArray_INT 5
#define PyArray_DOUBLE 12

/* array CONSTRUCTION is only used by attribute getters, which the oracle never calls */
static inline PyObject *PyArray_SimpleNewFromData(int nd, npy_intp *dims, int type, void *data) { (void)nd; (void)dims; (void)type; (void)data; return NULL; }
static inline PyObject *PyArray_Copy(PyArrayObject *a) { (void)a; return NULL; }
static inline PyObject *PyArray_Return(PyArrayObject *a) { return (PyObject *)a; }
static inline PyObject *PyArray_FromDimAndData(int nd, int *dims, int type, char *data) { (void)nd; (void)dims; (void)type; (void)data; return NULL; }

#endif
