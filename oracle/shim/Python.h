/* Minimal stand-in for the CPython-2 API, just enough to COMPILE the reference's neighbour-list / PROLSQ sources
 * (ensemble_hic/c/nblist.c, forcefield.c, prolsq.c, mathutils.c) as plain C where they lie -- test infrastructure
 * (oracle/), never linked into the product.  The reference extension targets Python 2 (Py_InitModule, PyInt_*,
 * tp_getattr, PyExc_StandardError) and cannot be built against CPython 3; its ARITHMETIC does not touch the
 * interpreter, so objects are plain heap structs here and the handful of API calls the sources make are reduced to
 * what the attribute setters need: an int and a double payload carried in every PyObject.
 * Written for this repository; nothing here is copied from CPython or from the reference. */
#ifndef EHIC_ORACLE_SHIM_PYTHON_H
#define EHIC_ORACLE_SHIM_PYTHON_H

#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>

struct _typeobject;

/* every object starts with this head; ival / dval are the shim's payload (PyInt_AsLong / PyFloat_AsDouble) */
#define PyObject_HEAD long ob_refcnt; struct _typeobject *ob_type; long shim_ival; double shim_dval;

typedef struct _object { PyObject_HEAD } PyObject;

typedef void (*destructor)(PyObject *);
typedef int (*printfunc)(PyObject *, FILE *, int);
typedef PyObject *(*getattrfunc)(PyObject *, char *);
typedef int (*setattrfunc)(PyObject *, char *, PyObject *);
typedef int (*cmpfunc)(PyObject *, PyObject *);
typedef PyObject *(*reprfunc)(PyObject *);
typedef long (*hashfunc)(PyObject *);
typedef PyObject *(*ternaryfunc)(PyObject *, PyObject *, PyObject *);
typedef PyObject *(*PyCFunction)(PyObject *, PyObject *);

/* field order = the positional initialisers of the reference's static type objects */
typedef struct _typeobject {
    PyObject_HEAD
    long ob_size;
    const char *tp_name;
    long tp_basicsize, tp_itemsize;
    destructor tp_dealloc;
    printfunc tp_print;
    getattrfunc tp_getattr;
    setattrfunc tp_setattr;
    cmpfunc tp_compare;
    reprfunc tp_repr;
    void *tp_as_number, *tp_as_sequence, *tp_as_mapping;
    hashfunc tp_hash;
    ternaryfunc tp_call;
    reprfunc tp_str;
    long shim_l0, shim_l1, shim_l2, shim_l3;
    const char *tp_doc;
} PyTypeObject;

#define PyObject_HEAD_INIT(type) 1, (struct _typeobject *)(type), 0, 0.0,

typedef struct { const char *ml_name; PyCFunction ml_meth; int ml_flags; } PyMethodDef;

extern PyObject shim_none, shim_exc;
extern char shim_last_error[256];
#define Py_None (&shim_none)
#define PyExc_StandardError (&shim_exc)
#define PyExc_ValueError (&shim_exc)
#define PyExc_TypeError (&shim_exc)
#define PyExc_AttributeError (&shim_exc)
#define Py_INCREF(o) ((void)(o))
#define Py_DECREF(o) ((void)(o))

static inline void PyErr_SetString(PyObject *exc, const char *msg)
{
    (void)exc;
    strncpy(shim_last_error, msg, sizeof(shim_last_error) - 1);
}
static inline void *shim_new(size_t size, PyTypeObject *type)
{
    PyObject *o = (PyObject *)calloc(1, size);
    if (o) { o->ob_refcnt = 1; o->ob_type = type; }
    return o;
}
#define PyObject_NEW(type, typeobj) ((type *)shim_new(sizeof(type), (typeobj)))
#define PyObject_Del(o) free(o)

static inline long PyInt_AsLong(PyObject *o) { return o->shim_ival; }
static inline double PyFloat_AsDouble(PyObject *o) { return o->shim_dval; }

/* the constructors parse an empty argument tuple; nothing else is called through this shim */
static inline int PyArg_ParseTuple(PyObject *args, const char *fmt, ...) { (void)args; return fmt[0] == '\0'; }
static inline PyObject *Py_BuildValue(const char *fmt, ...) { (void)fmt; return NULL; }
static inline PyObject *Py_FindMethod(PyMethodDef *m, PyObject *self, const char *name) { (void)m; (void)self; (void)name; return NULL; }

#endif
