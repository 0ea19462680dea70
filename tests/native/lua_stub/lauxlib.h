/* see lua.h in this directory: declarations only (Lua 5.2 reference manual, section 5) */
#ifndef LAUXLIB_STUB_H
#define LAUXLIB_STUB_H
#include "lua.h"
typedef struct luaL_Reg {
  const char *name;
  lua_CFunction func;
} luaL_Reg;
void luaL_checktype(lua_State *L, int arg, int t);
const char *luaL_optlstring(lua_State *L, int arg, const char *def, size_t *l);
#define luaL_optstring(L, n, d) (luaL_optlstring(L, (n), (d), NULL))
const char *luaL_checklstring(lua_State *L, int arg, size_t *l);
#define luaL_checkstring(L, n) (luaL_checklstring(L, (n), NULL))
lua_Integer luaL_optinteger(lua_State *L, int arg, lua_Integer def);
lua_Integer luaL_checkinteger(lua_State *L, int arg);
lua_Number luaL_optnumber(lua_State *L, int arg, lua_Number def);
void *luaL_checkudata(lua_State *L, int ud, const char *tname);
void luaL_setmetatable(lua_State *L, const char *tname);
int luaL_newmetatable(lua_State *L, const char *tname);
void luaL_setfuncs(lua_State *L, const luaL_Reg *l, int nup);
#define luaL_newlib(L, l) (lua_createtable(L, 0, (int)(sizeof(l) / sizeof((l)[0]) - 1)), luaL_setfuncs(L, l, 0))
#endif
