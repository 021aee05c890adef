// sqlite_dyn.h -- the handful of SQLite3 C API entry points the database layer needs, resolved at
// run time from libsqlite3.so.0.
//
// This image ships libsqlite3.so.0 (the library CPython's _sqlite3 links against) but no sqlite3.h
// and no libsqlite3.so development symlink, so the stable public C API is declared here by hand and
// bound with dlopen/dlsym.  Only documented, ABI-stable functions and constants are used.
#pragma once
#include <cstdint>

namespace b2mh {
namespace sq {

struct sqlite3;
struct sqlite3_stmt;

constexpr int kOk = 0;          // SQLITE_OK
constexpr int kRow = 100;       // SQLITE_ROW
constexpr int kDone = 101;      // SQLITE_DONE
constexpr int kOpenReadWrite = 0x2;
constexpr int kOpenCreate = 0x4;
constexpr int kTypeNull = 5;    // SQLITE_NULL
using Destructor = void (*)(void*);
inline Destructor Transient() { return reinterpret_cast<Destructor>(-1); }  // SQLITE_TRANSIENT

struct Api {
  int (*open_v2)(const char*, sqlite3**, int, const char*);
  int (*close)(sqlite3*);
  int (*exec)(sqlite3*, const char*, int (*)(void*, int, char**, char**), void*, char**);
  void (*free)(void*);
  int (*prepare_v2)(sqlite3*, const char*, int, sqlite3_stmt**, const char**);
  int (*step)(sqlite3_stmt*);
  int (*reset)(sqlite3_stmt*);
  int (*clear_bindings)(sqlite3_stmt*);
  int (*finalize)(sqlite3_stmt*);
  int (*bind_int64)(sqlite3_stmt*, int, int64_t);
  int (*bind_double)(sqlite3_stmt*, int, double);
  int (*bind_null)(sqlite3_stmt*, int);
  int (*bind_blob64)(sqlite3_stmt*, int, const void*, uint64_t, Destructor);
  int (*bind_text)(sqlite3_stmt*, int, const char*, int, Destructor);
  int64_t (*column_int64)(sqlite3_stmt*, int);
  double (*column_double)(sqlite3_stmt*, int);
  const void* (*column_blob)(sqlite3_stmt*, int);
  int (*column_bytes)(sqlite3_stmt*, int);
  const unsigned char* (*column_text)(sqlite3_stmt*, int);
  int (*column_type)(sqlite3_stmt*, int);
  const char* (*errmsg)(sqlite3*);
  int64_t (*last_insert_rowid)(sqlite3*);
  const char* (*libversion)();
};

// Loads libsqlite3 on first use; throws std::runtime_error when it (or a symbol) is missing.
const Api& api();

}  // namespace sq
}  // namespace b2mh
